// ros_stub: stand-ins for the ROS types that ransac_slam::System uses (include/ransac_slam/System.h:109-123, src/System.cpp).  Data members
// exist so that the reference's publisher code COMPILES; nothing is transported.  Used by (a) oracle/_ref/libref.so, which never
// compiles src/System.cpp (oracle/ref_driver.cpp replays its TrackRunning call sequence instead), and (b) tests/test_host_dropin.py,
// which compiles the reference's own src/System.cpp against the GPU-backed host classes of ransac_slam_b200/host/ to prove that they
// drop in.  TEST INFRASTRUCTURE.
#pragma once
#include <unistd.h>

#include <memory>
#include <string>
#include <vector>
namespace ros {
struct Time {
    static Time now() { return Time(); }
};
struct Duration {
    Duration(double = 0) {}
};
struct Rate {
    Rate(double) {}
    void sleep() {}
};
struct Publisher {
    template <class M> void publish(const M&) const {}
    int getNumSubscribers() const { return 1; }
};
struct NodeHandle {
    template <class M> Publisher advertise(const std::string&, int, bool = false) { return Publisher(); }
};
inline bool ok() { return true; }
inline void shutdown() {}
}  // namespace ros
#define ROS_WARN_ONCE(...) ((void)0)
namespace std_msgs {
struct Header {
    ros::Time stamp;
    std::string frame_id;
};
}  // namespace std_msgs
namespace geometry_msgs {
struct Point {
    double x = 0, y = 0, z = 0;
};
struct Vector3 {
    double x = 0, y = 0, z = 0;
};
struct Quaternion {
    double x = 0, y = 0, z = 0, w = 1;
};
struct Pose {
    Point position;
    Quaternion orientation;
};
struct PoseStamped {
    std_msgs::Header header;
    Pose pose;
};
struct PoseWithCovariance {
    Pose pose;
};
struct Twist {
    Vector3 linear, angular;
};
struct TwistWithCovariance {
    Twist twist;
};
struct Transform {
    Vector3 translation;
    Quaternion rotation;
};
struct TransformStamped {
    std_msgs::Header header;
    std::string child_frame_id;
    Transform transform;
};
}  // namespace geometry_msgs
namespace nav_msgs {
struct Path {
    std_msgs::Header header;
    std::vector<geometry_msgs::PoseStamped> poses;
};
struct Odometry {
    std_msgs::Header header;
    std::string child_frame_id;
    geometry_msgs::PoseWithCovariance pose;
    geometry_msgs::TwistWithCovariance twist;
};
}  // namespace nav_msgs
namespace std_msgs {
struct ColorRGBA {
    float r = 0, g = 0, b = 0, a = 0;
};
}  // namespace std_msgs
namespace visualization_msgs {
struct Marker {
    enum { ARROW = 0, CUBE = 1, SPHERE = 2, CYLINDER = 3, LINE_STRIP = 4, TEXT_VIEW_FACING = 9, ADD = 0 };
    std_msgs::Header header;
    std::string ns, text;
    int id = 0, type = 0, action = 0;
    geometry_msgs::Pose pose;
    geometry_msgs::Vector3 scale;
    std_msgs::ColorRGBA color;
    ros::Duration lifetime;
    std::vector<geometry_msgs::Point> points;
};
struct MarkerArray {
    std::vector<Marker> markers;
};
}  // namespace visualization_msgs
namespace sensor_msgs {
struct Image {};
typedef std::shared_ptr<Image> ImagePtr;
}  // namespace sensor_msgs
namespace cv { class Mat; }
namespace cv_bridge {
struct CvImage {
    CvImage(const std_msgs::Header&, const std::string&, const cv::Mat&) {}
    sensor_msgs::ImagePtr toImageMsg() const { return sensor_msgs::ImagePtr(); }
};
}  // namespace cv_bridge
namespace image_transport {
struct Publisher {
    template <class M> void publish(const M&) const {}
};
struct ImageTransport {
    ImageTransport() {}
    explicit ImageTransport(const ros::NodeHandle&) {}
    Publisher advertise(const std::string&, int) { return Publisher(); }
};
}  // namespace image_transport
namespace tf {
struct Quaternion {
    double x() const { return 0; }
    double y() const { return 0; }
    double z() const { return 0; }
    double w() const { return 1; }
};
struct Matrix3x3 {
    void setValue(double, double, double, double, double, double, double, double, double) {}
    void getRotation(Quaternion&) const {}
};
struct TransformBroadcaster {
    void sendTransform(const geometry_msgs::TransformStamped&) {}
};
}  // namespace tf

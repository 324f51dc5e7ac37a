// ros_stub: empty stand-ins for the ROS types that appear in the DECLARATION of ransac_slam::System (include/ransac_slam/System.h
// :109-123).  src/System.cpp (the ROS node: publishers, rviz markers) is not compiled; oracle/ref_driver.cpp replays its
// TrackRunning call sequence instead.  TEST INFRASTRUCTURE.
#pragma once
#include <string>
namespace ros {
struct NodeHandle {};
struct Publisher {};
struct Time {};
}  // namespace ros
namespace image_transport {
struct ImageTransport {};
struct Publisher {};
}  // namespace image_transport
namespace nav_msgs {
struct Path {};
struct Odometry {};
}  // namespace nav_msgs

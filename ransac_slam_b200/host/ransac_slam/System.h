// Host mirror of the reference's shared camera-parameter struct (include/ransac_slam/System.h:69-82).  The System class itself
// (ROS node, publishers) is out of scope; TrackRunning's call order is restated by host/replay_synth.cpp.
#pragma once
#include <string>

#include "../shim/linalg_shim.h"

namespace ransac_slam {
struct CamParam {
    double k1 = 0, k2 = 0;
    int nRows = 0, nCols = 0;
    double Cx = 0, Cy = 0, f = 0, dx = 0, dy = 0;
    std::string model;
    double K[9] = {0};  // row-major 3x3 (Eigen::Matrix3d in the reference)
};
// parses the Camera.* keys of the reference's OpenCV-YAML settings file (src/System.cpp:34-58)
bool load_camera_yaml(const std::string& path, CamParam* cam, int* min_features);
// tiny "key: value" reader for that YAML subset (cv::FileStorage is OpenCV C++, absent here)
bool yaml_get(const std::string& path, const std::string& key, double* out);
}  // namespace ransac_slam

// "ransac_slam/System.h" of the drop-in.
//   * Inside the reference's package: define RSLAM_REFERENCE_SYSTEM_H to the path of the reference's own
//     include/ransac_slam/System.h.  This header then forwards to it, so that the reference's System class (ROS node, publishers) and
//     its CamParam (include/ransac_slam/System.h:69-82) stay the reference's, while "ransac_slam/ExtendKF.h", "ransac_slam/Map.h" and
//     "ransac_slam/Tracking.h" -- which that header includes -- resolve to the GPU-backed classes of this directory.
//   * Stand-alone (the replay drivers of this repository): the shared camera-parameter struct alone; the System class is out of scope.
#pragma once
#include <string>

#include "../shim/linalg_shim.h"

#ifdef RSLAM_REFERENCE_SYSTEM_H
#include RSLAM_REFERENCE_SYSTEM_H
#else
namespace ransac_slam {
struct CamParam {
    double k1 = 0, k2 = 0;
    int nRows = 0, nCols = 0;
    double Cx = 0, Cy = 0, f = 0, dx = 0, dy = 0;
    std::string model;
    Eigen::Matrix3d K;
};
}  // namespace ransac_slam
#endif

namespace ransac_slam {
struct CamParam;
// parses the Camera.* keys of the reference's OpenCV-YAML settings file (src/System.cpp:34-58)
bool load_camera_yaml(const std::string& path, CamParam* cam, int* min_features);
// tiny "key: value" reader for that YAML subset (cv::FileStorage is OpenCV C++, absent in this repository's build image)
bool yaml_get(const std::string& path, const std::string& key, double* out);
}  // namespace ransac_slam

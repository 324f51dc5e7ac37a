// Host class with the reference's Map interface (include/ransac_slam/Map.h:16-63).  map_management runs ENTIRELY on the device
// (rslam_map_management: delete pass with the reference's run-ahead index, counters + flag reset, inverse-depth -> cartesian
// conversion, FAST-9 feature initialisation; src/Map.cpp:16-67), so x_k_k / p_k_k never leave HBM between frames.
#pragma once
#include <vector>

#include "ExtendKF.h"

namespace ransac_slam {
class ExtendKF;
struct Feature;
class Map {
  public:
    Map(const int min_fea, ExtendKF* m_ExtendKF);
    ~Map();
    void map_management(cv::Mat image, int step);  // src/Map.cpp:16-67
    // explicit uniform draws (2 per initialisation attempt, src/Map.cpp:231) for the next map_management() call; without them the draws
    // come from std::rand() like the reference (src/ExtendKF.cpp:230), mapped to [0,1)
    void set_uniform_draws(const double* u01, int n);
    // not in the reference: keep the map as it is (no delete / convert / initialise), only reset the per-frame flags -- for
    // synthetic sequences over a fixed map (configs C2-C5)
    void set_frozen(bool frozen) { frozen_ = frozen; }
    // {deleted, converted index or -1, initialised, attempts} of the last call
    const int* last_info() const { return info_; }

  private:
    int min_features;
    ExtendKF* mM_ExtendKF;
    std::vector<double> u01_;
    int info_[4] = {0, -1, 0, 0};
    bool frozen_ = false;
};
}  // namespace ransac_slam

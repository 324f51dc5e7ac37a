// Host class with the reference's Map interface (include/ransac_slam/Map.h:16-63).  On the measurement-update path only step 2 of
// map_management (counters + per-frame flag reset, src/Map.cpp:34-55) runs, on the device.  Feature deletion, inverse-depth ->
// cartesian conversion and FAST-9 initialisation are SURVEY 8(f) "next" rows and are not built yet: map_management leaves the
// map unchanged.
#pragma once
#include "ExtendKF.h"

namespace ransac_slam {
class Map {
  public:
    Map(const int min_fea, ExtendKF* m_ExtendKF);
    ~Map();
    void map_management(cv::Mat image, int step);

  private:
    int min_features;
    ExtendKF* mM_ExtendKF;
};
}  // namespace ransac_slam

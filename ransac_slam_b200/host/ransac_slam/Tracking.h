// Drop-in host class for the reference's Tracking (include/ransac_slam/Tracking.h:19-48).
#pragma once
#include <string>
#include <vector>

#include "ExtendKF.h"

namespace ransac_slam {
class Tracking {
  public:
    Tracking(const std::string& strSettingsFile, ExtendKF* m_ExtendKF);
    ~Tracking();
    void search_IC_matches(cv::Mat image);  // src/Tracking.cpp:32-70
    void ransac_hypotheses(void);           // src/Tracking.cpp:352-539
    void rescue_hi_inliers(void);           // src/Tracking.cpp:574-597
    // explicit uniform draws for the next ransac_hypotheses() call (north_star: hypothesis indices fed from the same seeded sequence); without
    // them the draws come from std::rand() like the reference (src/ExtendKF.cpp:230), mapped to [0,1)
    void set_uniform_draws(const double* u01, int n);
    rslam_ransac_result last_ransac() const { return last_; }

  private:
    ExtendKF* mT_ExtendKF;
    std::vector<double> u01_;
    rslam_ransac_result last_{};
};
}  // namespace ransac_slam

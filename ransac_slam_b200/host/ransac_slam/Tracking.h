// Drop-in host class for the reference's Tracking (include/ransac_slam/Tracking.h:19-48): same public methods.
#pragma once
#include <string>
#include <vector>

#include "ExtendKF.h"

namespace ransac_slam {
class ExtendKF;
class Tracking {
  public:
    Tracking(const std::string& strSettingsFile, ExtendKF* m_ExtendKF);
    ~Tracking();
    void search_IC_matches(cv::Mat image);  // src/Tracking.cpp:32-70
    // The stages of search_IC_matches one by one (public in the reference, src/Tracking.cpp:540-573, 164-278, 279-351).  On the device h_i,
    // H_i and S_i are produced together by one kernel and the patch warp by another, for ALL features of the filter:
    //   calculate_derivatives : measurement prediction + Jacobians + S_i at the prior (the argument must be x_k_km1, as at every call site)
    //   pred_patch_fc         : the warp of every predicted feature (the per-feature arguments are ignored: it is a batch operation)
    //   matching              : the ZNCC search on `image` with the current predictions
    void calculate_derivatives(Eigen::VectorXd xk_km1);
    void calculate_Hi_cartesian(Eigen::VectorXd x_v, Eigen::VectorXd yi, int order, Eigen::MatrixXd& Hi);      // dense 2 x n from the device's sparse H_i
    void calculate_Hi_inverse_depth(Eigen::VectorXd x_v, Eigen::VectorXd yi, int order, Eigen::MatrixXd& Hi);  // (after calculate_derivatives)
    void pred_patch_fc(int order, Eigen::Vector3d XYZ_w);
    void matching(cv::Mat image);
    void ransac_hypotheses(void);   // src/Tracking.cpp:352-539
    void rescue_hi_inliers(void);   // src/Tracking.cpp:574-597
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

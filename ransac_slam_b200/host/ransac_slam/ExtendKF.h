// Drop-in host class for the reference's ExtendKF (include/ransac_slam/ExtendKF.h:44-180) on the measurement-update path.
// Same constructor signature, same public data members and the same method names; the arithmetic runs on the GPU through the
// C ABI (include/rslam.h).  The authoritative state lives in HBM; the public members below are host mirrors refreshed by
// sync_to_host() and pushed by sync_to_device() (the reference mutates them in place from Map / Tracking).
#pragma once
#include <string>
#include <vector>

#include "../../../include/rslam.h"
#include "System.h"

namespace ransac_slam {

struct Feature {  // include/ransac_slam/ExtendKF.h:14-42 (fields the path reads or writes)
    Eigen::MatrixXd patch_when_initialized;  // 41 x 41
    Eigen::MatrixXd patch_when_matching;     // 13 x 13
    double r_wc_when_initialized[3] = {0, 0, 0};
    double R_wc_when_initialized[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double uv_when_initialized[2] = {0, 0};
    int half_patch_size_when_initialized = 20;
    int half_patch_size_when_matching = 6;
    int times_predicted = 0, times_measured = 0;
    long long init_frame = 0;
    std::string type = "inversedepth";
    bool individually_compatible = false, low_innovation_inlier = false, high_innovation_inlier = false;
    Eigen::VectorXd z;   // size 2 when matched, 0 otherwise
    Eigen::VectorXd h;   // size 2 when predicted, 0 otherwise
    Eigen::MatrixXd H;   // dense 2 x n, materialised on demand (materialize_H)
    Eigen::MatrixXd S;   // 2 x 2
    double Hc[14] = {0}, Hf[12] = {0};  // the structurally non-zero part kept by the device
};

class ExtendKF {
  public:
    ExtendKF(const std::string& strSettingsFile, CamParam* param, std::string type);
    ~ExtendKF();

    void initialize_x_and_p(void);                       // src/ExtendKF.cpp:32-55
    void ekf_prediction(void);                           // src/ExtendKF.cpp:333-388
    void ekf_update_li_inliers(void);                    // src/ExtendKF.cpp:559-596
    void ekf_update_hi_inliers(void);                    // src/ExtendKF.cpp:640-678
    void predict_camera_measurements(Eigen::VectorXd xkk);  // src/ExtendKF.cpp:56-90 (device: runs with the search / rescue stage)

    // --- device residency -------------------------------------------------------------------------------------------------
    // push x_k_k, p_k_k, feature types and predicted patches to the GPU (after Map added/removed features)
    int sync_to_device();
    // pull x_k_k, p_k_k (want_P), per-feature h/S/z/flags into the public members
    int sync_to_host(bool want_P = true);
    void materialize_H(int feature);  // fills features_info[i].H (2 x n) from the sparse device form
    rslam_filter* device_handle() { return dev_; }
    int last_status() const { return status_; }

    std::vector<Feature> features_info;
    CamParam* cam;
    double v_0 = 0, std_v_0 = 0.025, w_0 = 1e-11, std_w_0 = 0.025;
    Eigen::VectorXd x_k_k;
    Eigen::MatrixXd p_k_k;
    double std_a = 0.007, std_alpha = 0.007, std_z = 1.0;
    Eigen::VectorXd x_k_km1;
    Eigen::MatrixXd p_k_km1;  // not mirrored: the device keeps one covariance, updated in place (see include/rslam.h)

  private:
    friend class Tracking;
    friend class Map;
    int ensure_device(int max_features);
    std::string filter_type;
    rslam_filter* dev_ = nullptr;
    int dev_capacity_ = 0;
    int status_ = 0;
};

}  // namespace ransac_slam

// Drop-in host class for the reference's ExtendKF (include/ransac_slam/ExtendKF.h:44-180).
// Same constructor signature, same public data members, the same method names and signatures.  The measurement-update path runs on
// the GPU through the C ABI (include/rslam.h): the authoritative state lives in HBM; the public members below are host mirrors
// refreshed by sync_to_host() and pushed by sync_to_device() (the reference mutates them in place from Map / Tracking).  The
// camera-model helpers other code calls (src/Map.cpp:141,222,234,276,291,347,350,362,377; src/System.cpp:314,403) are O(1) per
// point and stay on the host, formula by formula.
#pragma once
#include <string>
#include <vector>

#include "../../../include/rslam.h"
#include "System.h"

namespace ransac_slam {
struct CamParam;

struct Feature {  // include/ransac_slam/ExtendKF.h:14-42, member for member
    Eigen::MatrixXd patch_when_initialized;  // 41 x 41
    Eigen::MatrixXd patch_when_matching;     // 13 x 13
    Eigen::Vector3d r_wc_when_initialized;
    Eigen::Matrix3d R_wc_when_initialized;
    Eigen::RowVector2d uv_when_initialized;
    int half_patch_size_when_initialized = 20;
    int half_patch_size_when_matching = 6;
    int times_predicted = 0, times_measured = 0;
    long long int init_frame = 0;
    Eigen::Vector2d init_measurement;
    std::string type = "inversedepth";
    Eigen::VectorXd yi;
    bool individually_compatible = false, low_innovation_inlier = false, high_innovation_inlier = false;
    Eigen::VectorXd z;     // size 2 when matched, 0 otherwise
    Eigen::RowVectorXd h;  // size 2 when predicted, 0 otherwise
    Eigen::MatrixXd H;     // dense 2 x n, materialised on demand (ExtendKF::materialize_H): the device keeps the non-zero part
    Eigen::MatrixXd S;     // 2 x 2
    int state_size = 6, measurement_size = 2;
    Eigen::Matrix2d R;
    // not in the reference: the structurally non-zero part of H_i as the device stores it (d h / d (r, q) 2 x 7, d h / d y_i 2 x 6)
    double Hc[14] = {0}, Hf[12] = {0};
};

class ExtendKF {
  public:
    ExtendKF(const std::string& strSettingsFile, CamParam* param, std::string type);
    ~ExtendKF();

    void initialize_x_and_p(void);                          // src/ExtendKF.cpp:32-55
    void predict_camera_measurements(Eigen::VectorXd xkk);  // src/ExtendKF.cpp:56-90 (device: runs with the search / rescue stage)
    void ekf_prediction(void);                              // src/ExtendKF.cpp:333-388
    void ekf_update_li_inliers(void);                       // src/ExtendKF.cpp:559-596
    void ekf_update_hi_inliers(void);                       // src/ExtendKF.cpp:640-678
    // src/ExtendKF.cpp:597-639 on the device.  H must have the structure of stacked measurement Jacobians (per row pair: columns 0..6 and
    // ONE feature block non-zero), R must be the identity (what every call site passes, src/ExtendKF.cpp:594,676); anything else sets
    // last_status() = RSLAM_ERR_INVALID and leaves x_k_k / p_k_k alone.  Writes x_k_k and p_k_k like the reference.
    void update(Eigen::VectorXd x_km_k, Eigen::MatrixXd p_km_k, Eigen::MatrixXd H, Eigen::MatrixXd R, Eigen::VectorXd z, Eigen::VectorXd h);

    // camera model and small Jacobians, host side (file:line of the reference next to each definition in host_classes.cpp)
    Eigen::Matrix3d q2r(Eigen::VectorXd q_in);
    void hi_cartesian(Eigen::Vector3d hrl, Eigen::MatrixXd& zi);
    void hi_inverse_depth(Eigen::Vector3d hrl, Eigen::MatrixXd& zi);
    Eigen::Vector3d inversedepth2cartesian(Eigen::VectorXd inverse_depth);
    Eigen::Vector2d hu(Eigen::Vector3d yi);
    void distort_fm(Eigen::MatrixXd uv, Eigen::MatrixXd& uvd);
    void undistort_fm(Eigen::MatrixXd uvd, Eigen::MatrixXd& uvu);
    double RandomGenerator(const int low, const int high);
    Eigen::MatrixXd rand(int row, int column, double min, double max);
    void hinv(Eigen::VectorXd uvd, Eigen::VectorXd Xv, double initial_rho, Eigen::VectorXd& newFeature);
    Eigen::Matrix<double, 3, 4> dRq_times_a_by_dq(Eigen::VectorXd q, Eigen::Vector3d aMat);
    Eigen::Matrix2d jacob_undistor_fm(Eigen::VectorXd uvd);

    // --- device residency (not in the reference) ----------------------------------------------------------------------------
    // push x_k_k, p_k_k, feature types and predicted patches to the GPU (after host code added / removed features)
    int sync_to_device();
    // pull x_k_k, p_k_k (want_P), per-feature h/S/z/flags into the public members; want_prior_P also fills p_k_km1 with the device's
    // covariance (meaningful between ekf_prediction and the first update: the device keeps ONE covariance, updated in place)
    int sync_to_host(bool want_P = true, bool want_prior_P = false);
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
    Eigen::MatrixXd p_k_km1;

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

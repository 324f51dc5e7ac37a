// host_classes.cpp -- ExtendKF / Tracking / Map with the reference's interfaces, delegating the measurement-update path to the
// C ABI (include/rslam.h).  No arithmetic of the path runs on the CPU here; failures set last_status() (the reference's own
// failure mode is exit(-1) / Eigen asserts, which a library must not copy).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

#include "ransac_slam/ExtendKF.h"
#include "ransac_slam/Map.h"
#include "ransac_slam/Tracking.h"

namespace ransac_slam {

bool yaml_get(const std::string& path, const std::string& key, double* out) {
    std::ifstream in(path);
    if (!in) return false;
    std::string line;
    while (std::getline(in, line)) {
        size_t hash = line.find('#');
        if (hash != std::string::npos) line = line.substr(0, hash);
        size_t colon = line.find(':');
        if (colon == std::string::npos) continue;
        std::string k = line.substr(0, colon);
        size_t a = k.find_first_not_of(" \t"), b = k.find_last_not_of(" \t");
        if (a == std::string::npos) continue;
        k = k.substr(a, b - a + 1);
        if (k != key) continue;
        std::istringstream v(line.substr(colon + 1));
        double d;
        if (v >> d) {
            *out = d;
            return true;
        }
    }
    return false;
}

bool load_camera_yaml(const std::string& path, CamParam* cam, int* min_features) {  // src/System.cpp:34-60
    double d = 0, cxd = 0, cyd = 0, v = 0;
    bool ok = yaml_get(path, "Camera.k1", &cam->k1) && yaml_get(path, "Camera.k2", &cam->k2);
    ok = ok && yaml_get(path, "Camera.nRows", &v);
    cam->nRows = (int)v;
    ok = ok && yaml_get(path, "Camera.nCols", &v);
    cam->nCols = (int)v;
    ok = ok && yaml_get(path, "Camera.d", &d) && yaml_get(path, "Camera.cx_d", &cxd) && yaml_get(path, "Camera.cy_d", &cyd);
    ok = ok && yaml_get(path, "Camera.fps", &cam->f);  // "fps" holds the focal length in mm (src/System.cpp:49)
    ok = ok && yaml_get(path, "Camera.dx", &cam->dx) && yaml_get(path, "Camera.dy", &cam->dy);
    if (!ok) return false;
    cam->Cx = cxd / d;
    cam->Cy = cyd / d;
    const double K[9] = {cam->f / d, 0, cam->Cx, 0, cam->f / d, cam->Cy, 0, 0, 1};
    std::memcpy(cam->K, K, sizeof(K));
    if (min_features && yaml_get(path, "min_number_of_features_in_image", &v)) *min_features = (int)v;
    return true;
}

// ---------------------------------------------------------------------------------------------------------------------
ExtendKF::ExtendKF(const std::string& strSettingsFile, CamParam* param, std::string type) : cam(param), filter_type(type) {
    // src/ExtendKF.cpp:16-23 ; a missing file keeps the defaults of initialize_param.yaml instead of exit(-1)
    if (!strSettingsFile.empty()) {
        yaml_get(strSettingsFile, "Sigma.a", &std_a);
        yaml_get(strSettingsFile, "Sigma.alpha", &std_alpha);
        yaml_get(strSettingsFile, "Sigma.noise", &std_z);
        yaml_get(strSettingsFile, "Velocity.v0", &v_0);
        yaml_get(strSettingsFile, "Velocity.stdv0", &std_v_0);
        yaml_get(strSettingsFile, "Velocity.w0", &w_0);
        yaml_get(strSettingsFile, "Velocity.stdw0", &std_w_0);
    }
}
ExtendKF::~ExtendKF() {
    if (dev_) rslam_destroy(dev_);
}

void ExtendKF::initialize_x_and_p(void) {  // src/ExtendKF.cpp:32-55 (P(5,5) left 0: quirk Q7)
    const double eps = 2.220446049250313e-16;
    x_k_k.resize(13);
    x_k_k(3) = 1;
    for (int i = 7; i < 10; i++) x_k_k(i) = v_0;
    for (int i = 10; i < 13; i++) x_k_k(i) = w_0;
    p_k_k.resize(13, 13);
    for (int i : {0, 1, 2, 3, 4, 6}) p_k_k(i, i) = eps;
    for (int i = 7; i < 10; i++) p_k_k(i, i) = std_v_0 * std_v_0;
    for (int i = 10; i < 13; i++) p_k_k(i, i) = std_w_0 * std_w_0;
}

int ExtendKF::ensure_device(int max_features) {
    if (dev_ && max_features <= dev_capacity_) return 0;
    if (dev_) rslam_destroy(dev_);
    dev_ = nullptr;
    rslam_camera c{cam->k1, cam->k2, cam->nRows, cam->nCols, cam->Cx, cam->Cy, cam->f, cam->dx, cam->dy};
    rslam_params p;
    rslam_default_params(&p);
    p.std_a = std_a;
    p.std_alpha = std_alpha;
    p.std_z = std_z;
    int cap = max_features < 256 ? 256 : max_features + max_features / 2;  // the map grows on the device (Map::map_management)
    status_ = rslam_create(&c, &p, cap, 1, 0, &dev_);
    if (status_ == 0) dev_capacity_ = cap;
    return status_;
}

int ExtendKF::sync_to_device() {
    const int N = (int)features_info.size();
    if ((status_ = ensure_device(N))) return status_;
    std::vector<int> types(N + 1);
    std::vector<double> patches((size_t)N * 169 + 1, 0.0);
    for (int i = 0; i < N; i++) {
        types[i] = features_info[i].type == "cartesian" ? 1 : 0;
        const Eigen::MatrixXd& pm = features_info[i].patch_when_matching;
        if (pm.rows() == 13 && pm.cols() == 13)
            for (int r = 0; r < 13; r++)
                for (int c = 0; c < 13; c++) patches[(size_t)i * 169 + r * 13 + c] = pm(r, c);
    }
    const int n = x_k_k.rows();
    status_ = rslam_upload_state(dev_, 0, 0, x_k_k.data(), p_k_k.data(), n, p_k_k.rows(), types.data(), N);
    if (status_ == 0 && N) status_ = rslam_upload_patches(dev_, 0, patches.data(), N);
    // when every feature carries its 41 x 41 initialisation patch (Map::initialize_a_features, src/Map.cpp:286-294) the predicted
    // appearance is warped on the device (Tracking::pred_patch_fc); otherwise patch_when_matching is used as given
    // (an empty map counts as "has them": its features will be created on the device by Map::map_management, patches included)
    bool have_init = true;
    for (int i = 0; i < N && have_init; i++)
        have_init = features_info[i].patch_when_initialized.rows() == 41 && features_info[i].patch_when_initialized.cols() == 41;
    if (status_ == 0 && have_init && N > 0) {
        std::vector<uint8_t> p41((size_t)N * 1681);
        std::vector<double> r(3 * (size_t)N), R(9 * (size_t)N), uv(2 * (size_t)N);
        for (int i = 0; i < N; i++) {
            const Feature& ft = features_info[i];
            for (int a = 0; a < 41; a++)
                for (int b = 0; b < 41; b++) p41[(size_t)i * 1681 + a * 41 + b] = (uint8_t)ft.patch_when_initialized(a, b);
            for (int k = 0; k < 3; k++) r[3 * i + k] = ft.r_wc_when_initialized[k];
            for (int k = 0; k < 9; k++) R[9 * i + k] = ft.R_wc_when_initialized[k];
            uv[2 * i] = ft.uv_when_initialized[0];
            uv[2 * i + 1] = ft.uv_when_initialized[1];
        }
        status_ = rslam_upload_feature_init(dev_, 0, p41.data(), r.data(), R.data(), uv.data(), N);
    }
    if (status_ == 0) status_ = rslam_set_patch_warp(dev_, have_init ? 1 : 0);
    return status_;
}

int ExtendKF::sync_to_host(bool want_P) {
    if (!dev_) return status_ = RSLAM_ERR_INVALID;
    const int n = rslam_state_dim(dev_, 0), N = rslam_num_features(dev_, 0);
    x_k_k.resize(n);
    x_k_km1.resize(n);
    if (want_P) p_k_k.resize(n, n);
    if ((status_ = rslam_download_state(dev_, 0, 0, x_k_k.data(), want_P ? p_k_k.data() : nullptr, n))) return status_;
    if ((status_ = rslam_download_state(dev_, 0, 1, x_k_km1.data(), nullptr, n))) return status_;
    std::vector<double> h(2 * N), S(4 * N), z(2 * N), Hc(14 * N), Hf(12 * N);
    std::vector<uint8_t> fl(4 * N);
    std::vector<int> cnt(2 * N);
    if ((status_ = rslam_download_features(dev_, 0, h.data(), S.data(), z.data(), fl.data(), cnt.data()))) return status_;
    if ((status_ = rslam_download_H(dev_, 0, Hc.data(), Hf.data()))) return status_;
    // the map may have changed on the device (delete / convert / initialise): the mirror follows its size and types
    std::vector<int> types(N);
    if (N && (status_ = rslam_feature_types(dev_, 0, types.data()))) return status_;
    features_info.resize(N);
    for (int i = 0; i < N; i++) {
        Feature& f = features_info[i];
        f.type = types[i] ? "cartesian" : "inversedepth";
        f.individually_compatible = fl[4 * i + 1];
        f.low_innovation_inlier = fl[4 * i + 2];
        f.high_innovation_inlier = fl[4 * i + 3];
        f.times_predicted = cnt[2 * i];
        f.times_measured = cnt[2 * i + 1];
        if (fl[4 * i]) {
            f.h.resize(2);
            f.h(0) = h[2 * i];
            f.h(1) = h[2 * i + 1];
            f.S.resize(2, 2);
            for (int a = 0; a < 2; a++)
                for (int b = 0; b < 2; b++) f.S(a, b) = S[4 * i + 2 * a + b];
        } else {
            f.h.resize(0);
            f.S.resize(0, 0);
        }
        if (f.individually_compatible) {
            f.z.resize(2);
            f.z(0) = z[2 * i];
            f.z(1) = z[2 * i + 1];
        } else {
            f.z.resize(0);
        }
        std::memcpy(f.Hc, &Hc[14 * i], sizeof(f.Hc));
        std::memcpy(f.Hf, &Hf[12 * i], sizeof(f.Hf));
    }
    return status_;
}

void ExtendKF::materialize_H(int idx) {  // dense 2 x n row pair as the reference stores it (src/Tracking.cpp:128-129)
    const int n = x_k_k.rows();
    Feature& f = features_info[idx];
    f.H.resize(2, n);
    int off = 13;
    for (int i = 0; i < idx; i++) off += features_info[i].type == "cartesian" ? 3 : 6;
    const int fs = f.type == "cartesian" ? 3 : 6;
    for (int a = 0; a < 2; a++) {
        for (int c = 0; c < 7; c++) f.H(a, c) = f.Hc[a * 7 + c];
        for (int c = 0; c < fs; c++) f.H(a, off + c) = f.Hf[a * 6 + c];
    }
}

void ExtendKF::ekf_prediction(void) { status_ = dev_ ? rslam_ekf_prediction(dev_) : RSLAM_ERR_INVALID; }
void ExtendKF::ekf_update_li_inliers(void) { status_ = dev_ ? rslam_update_li(dev_) : RSLAM_ERR_INVALID; }
void ExtendKF::ekf_update_hi_inliers(void) { status_ = dev_ ? rslam_update_hi(dev_) : RSLAM_ERR_INVALID; }
void ExtendKF::predict_camera_measurements(Eigen::VectorXd) {
    // on the device h_i is produced together with H_i and S_i by the search / rescue stage (rslam_search_ic_matches,
    // rslam_rescue_hi); a stand-alone call has nothing left to do on the measurement-update path
}

// ---------------------------------------------------------------------------------------------------------------------
Tracking::Tracking(const std::string&, ExtendKF* m_ExtendKF) : mT_ExtendKF(m_ExtendKF) {}
Tracking::~Tracking() {}
void Tracking::set_uniform_draws(const double* u01, int n) { u01_.assign(u01, u01 + n); }

void Tracking::search_IC_matches(cv::Mat image) {
    rslam_filter* d = mT_ExtendKF->dev_;
    if (!d) {
        mT_ExtendKF->status_ = RSLAM_ERR_INVALID;
        return;
    }
    int rc = rslam_set_image(d, 0, image.data, image.rows, image.cols, (int)image.step, 0);
    if (rc == 0) rc = rslam_search_ic_matches(d);
    mT_ExtendKF->status_ = rc;
}
void Tracking::ransac_hypotheses(void) {
    rslam_filter* d = mT_ExtendKF->dev_;
    if (!d) {
        mT_ExtendKF->status_ = RSLAM_ERR_INVALID;
        return;
    }
    if (u01_.empty()) {  // reference behaviour: libc rand (src/ExtendKF.cpp:230), kept in [0,1) (quirk Q8)
        u01_.resize(1000);
        for (double& u : u01_) u = (double)std::rand() / ((double)RAND_MAX + 1.0);
    }
    int rc = rslam_ransac_hypotheses(d, u01_.data(), (int)u01_.size());
    if (rc == 0) rc = rslam_ransac_result_get(d, 0, &last_);
    mT_ExtendKF->status_ = rc;
    u01_.clear();
}
void Tracking::rescue_hi_inliers(void) {
    rslam_filter* d = mT_ExtendKF->dev_;
    mT_ExtendKF->status_ = d ? rslam_rescue_hi(d) : RSLAM_ERR_INVALID;
}

// ---------------------------------------------------------------------------------------------------------------------
Map::Map(const int min_fea, ExtendKF* m_ExtendKF) : min_features(min_fea), mM_ExtendKF(m_ExtendKF) {}
Map::~Map() {}
void Map::set_uniform_draws(const double* u01, int n) { u01_.assign(u01, u01 + n); }
void Map::map_management(cv::Mat image, int step) {  // src/Map.cpp:16-67, on the device
    ExtendKF* kf = mM_ExtendKF;
    if (!kf->dev_ && kf->sync_to_device()) return;  // first frame: push the freshly initialised 13-state filter
    rslam_filter* d = kf->dev_;
    if (frozen_) {  // fixed synthetic maps: only step 2, the counters + flag reset (src/Map.cpp:34-55)
        kf->status_ = rslam_begin_frame(d);
        return;
    }
    if (u01_.empty()) {  // reference behaviour: ExtendKF::rand(2,1,0,1) per attempt, at most 50 attempts (src/Map.cpp:200,231)
        u01_.resize(100);
        for (double& u : u01_) u = (double)std::rand() / ((double)RAND_MAX + 1.0);
    }
    int rc = rslam_set_image(d, 0, image.data, image.rows, image.cols, (int)image.step, 0);
    if (rc == 0) rc = rslam_map_management(d, 0, step, min_features, 1, u01_.data(), (int)(u01_.size() / 2), info_);
    kf->status_ = rc;
    u01_.clear();
}

}  // namespace ransac_slam

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
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) cam->K(i, j) = K[3 * i + j];
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
            for (int k = 0; k < 3; k++) r[3 * i + k] = ft.r_wc_when_initialized(k);
            for (int k = 0; k < 9; k++) R[9 * i + k] = ft.R_wc_when_initialized(k / 3, k % 3);
            uv[2 * i] = ft.uv_when_initialized(0);
            uv[2 * i + 1] = ft.uv_when_initialized(1);
        }
        status_ = rslam_upload_feature_init(dev_, 0, p41.data(), r.data(), R.data(), uv.data(), N);
    }
    if (status_ == 0) status_ = rslam_set_patch_warp(dev_, have_init ? 1 : 0);
    return status_;
}

int ExtendKF::sync_to_host(bool want_P, bool want_prior_P) {
    if (!dev_) return status_ = RSLAM_ERR_INVALID;
    const int n = rslam_state_dim(dev_, 0), N = rslam_num_features(dev_, 0);
    x_k_k.resize(n);
    x_k_km1.resize(n);
    if (want_P) p_k_k.resize(n, n);
    if (want_prior_P) p_k_km1.resize(n, n);
    if ((status_ = rslam_download_state(dev_, 0, 0, x_k_k.data(), want_P ? p_k_k.data() : nullptr, n))) return status_;
    if ((status_ = rslam_download_state(dev_, 0, 1, x_k_km1.data(), want_prior_P ? p_k_km1.data() : nullptr, n))) return status_;
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
    for (int a = 0; a < 2; a++)
        for (int c = 0; c < n; c++) f.H(a, c) = 0.0;
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

// ---- ExtendKF::update (src/ExtendKF.cpp:597-639) through the device ----------------------------------------------------------------
void ExtendKF::update(Eigen::VectorXd x_km_k, Eigen::MatrixXd p_km_k, Eigen::MatrixXd H, Eigen::MatrixXd R, Eigen::VectorXd z, Eigen::VectorXd h) {
    const int n = (int)x_km_k.rows(), k = (int)z.rows(), N = (int)features_info.size();
    if (k == 0) {  // :635-638
        x_k_k = x_km_k;
        p_k_k = p_km_k;
        status_ = 0;
        return;
    }
    status_ = RSLAM_ERR_INVALID;
    if (k % 2 || H.rows() != k || H.cols() != n || h.rows() != k || R.rows() != k || R.cols() != k || p_km_k.rows() != n) return;
    for (int i = 0; i < k; i++)
        for (int j = 0; j < k; j++)
            if (R(i, j) != (i == j ? 1.0 : 0.0)) return;
    std::vector<int> types(N + 1), offs(N + 1);
    int off = 13;
    for (int i = 0; i < N; i++) {
        types[i] = features_info[i].type == "cartesian" ? 1 : 0;
        offs[i] = off;
        off += types[i] ? 3 : 6;
    }
    if (off != n) return;
    std::vector<double> hh(2 * (size_t)N + 1, 0.0), zz(2 * (size_t)N + 1, 0.0), Hc(14 * (size_t)N + 1, 0.0), Hf(12 * (size_t)N + 1, 0.0);
    std::vector<uint8_t> fl(4 * (size_t)N + 1, 0);
    int last = -1;
    for (int t = 0; t < k / 2; t++) {
        int feat = -1;
        for (int c = 7; c < n; c++) {
            if (H(2 * t, c) == 0.0 && H(2 * t + 1, c) == 0.0) continue;
            if (c < 13) return;  // velocity columns must be structurally zero
            int i = 0;
            while (i + 1 < N && offs[i + 1] <= c) i++;
            if (feat >= 0 && feat != i) return;  // more than one feature block in a row pair
            feat = i;
        }
        if (feat < 0 || feat <= last) return;  // the device stacks the measurements in feature order
        last = feat;
        const int fs = types[feat] ? 3 : 6;
        for (int a = 0; a < 2; a++) {
            for (int c = 0; c < 7; c++) Hc[14 * feat + 7 * a + c] = H(2 * t + a, c);
            for (int c = 0; c < fs; c++) Hf[12 * feat + 6 * a + c] = H(2 * t + a, offs[feat] + c);
            hh[2 * feat + a] = h(2 * t + a);
            zz[2 * feat + a] = z(2 * t + a);
        }
        fl[4 * feat] = fl[4 * feat + 1] = fl[4 * feat + 2] = 1;  // has_h, individually_compatible, low_innovation_inlier
    }
    if ((status_ = ensure_device(N))) return;
    if ((status_ = rslam_upload_state(dev_, 0, 1, x_km_k.data(), p_km_k.data(), n, n, types.data(), N))) return;
    if ((status_ = rslam_upload_linearisation(dev_, 0, hh.data(), Hc.data(), Hf.data(), zz.data(), fl.data()))) return;
    if ((status_ = rslam_update_li(dev_))) return;  // prior = x_k_km1, writes x_k_k and the covariance in place
    x_k_k.resize(n);
    p_k_k.resize(n, n);
    status_ = rslam_download_state(dev_, 0, 0, x_k_k.data(), p_k_k.data(), n);
}

// ---- camera model and small Jacobians on the host (O(1) per point; formulas of the reference, file:line) ------------------------------
Eigen::Matrix3d ExtendKF::q2r(Eigen::VectorXd q) {  // src/ExtendKF.cpp:91-102
    const double r = q(0), x = q(1), y = q(2), z = q(3);
    Eigen::Matrix3d R;
    R(0, 0) = r * r + x * x - y * y - z * z;
    R(0, 1) = 2 * (x * y - r * z);
    R(0, 2) = 2 * (z * x + r * y);
    R(1, 0) = 2 * (x * y + r * z);
    R(1, 1) = r * r - x * x + y * y - z * z;
    R(1, 2) = 2 * (y * z - r * x);
    R(2, 0) = 2 * (z * x - r * y);
    R(2, 1) = 2 * (y * z + r * x);
    R(2, 2) = r * r - x * x - y * y + z * z;
    return R;
}
Eigen::Vector2d ExtendKF::hu(Eigen::Vector3d yi) {  // src/ExtendKF.cpp:153-174
    Eigen::Vector2d uv;
    uv(0) = cam->Cx + (yi(0) / yi(2)) * cam->f * (1.0 / cam->dx);
    uv(1) = cam->Cy + (yi(1) / yi(2)) * cam->f * (1.0 / cam->dy);
    return uv;
}
void ExtendKF::distort_fm(Eigen::MatrixXd uv, Eigen::MatrixXd& uvd) {  // src/ExtendKF.cpp:175-204, 2 x m, ten Newton steps
    const int m = (int)uv.cols();
    Eigen::MatrixXd out(2, m);
    const double k1 = cam->k1, k2 = cam->k2;
    for (int c = 0; c < m; c++) {
        const double xu = (uv(0, c) - cam->Cx) * cam->dx, yu = (uv(1, c) - cam->Cy) * cam->dy;
        const double ru = std::sqrt(std::pow(xu, 2) + std::pow(yu, 2));
        double rd = ru / (1 + k1 * std::pow(ru, 2) + k2 * std::pow(ru, 4));
        for (int it = 0; it < 10; it++) {
            const double f = rd + k1 * std::pow(rd, 3) + k2 * std::pow(rd, 5) - ru;
            const double fp = 1 + 3 * k1 * std::pow(rd, 2) + 5 * k2 * std::pow(rd, 4);
            rd = rd - f / fp;
        }
        const double D = 1 + k1 * std::pow(rd, 2) + k2 * std::pow(rd, 4);
        out(0, c) = xu / D / cam->dx + cam->Cx;
        out(1, c) = yu / D / cam->dy + cam->Cy;
    }
    uvd = out;
}
void ExtendKF::undistort_fm(Eigen::MatrixXd uvd, Eigen::MatrixXd& uvu) {  // src/ExtendKF.cpp:266-285, 2 x m
    const int m = (int)uvd.cols();
    Eigen::MatrixXd out(2, m);
    for (int c = 0; c < m; c++) {
        const double xd = (uvd(0, c) - cam->Cx) * cam->dx, yd = (uvd(1, c) - cam->Cy) * cam->dy;
        const double rd = std::sqrt(std::pow(xd, 2) + std::pow(yd, 2));
        const double D = 1 + cam->k1 * std::pow(rd, 2) + cam->k2 * std::pow(rd, 4);
        out(0, c) = xd * D / cam->dx + cam->Cx;
        out(1, c) = yd * D / cam->dy + cam->Cy;
    }
    uvu = out;
}
void ExtendKF::hi_cartesian(Eigen::Vector3d hrl, Eigen::MatrixXd& zi) {  // src/ExtendKF.cpp:103-132: +-60 degree gate, image gate
    const double ax = std::atan2(hrl(0), hrl(2)) * 180 / M_PI, ay = std::atan2(hrl(1), hrl(2)) * 180 / M_PI;
    if (ax < -60 || ax > 60 || ay < -60 || ay > 60) {
        zi.resize(0, 0);
        return;
    }
    Eigen::Vector2d u = hu(hrl);
    Eigen::MatrixXd uu(2, 1), ud;
    uu(0, 0) = u(0);
    uu(1, 0) = u(1);
    distort_fm(uu, ud);
    if (ud(0, 0) > 0 && ud(0, 0) < cam->nCols && ud(1, 0) > 0 && ud(1, 0) < cam->nRows)
        zi = ud;
    else
        zi.resize(0, 0);
}
void ExtendKF::hi_inverse_depth(Eigen::Vector3d, Eigen::MatrixXd&) {}  // empty in the reference too (src/ExtendKF.cpp:133-136)
Eigen::Vector3d ExtendKF::inversedepth2cartesian(Eigen::VectorXd y) {  // src/ExtendKF.cpp:137-152
    const double th = y(3), ph = y(4), rho = y(5);
    const double m[3] = {std::cos(ph) * std::sin(th), -std::sin(ph), std::cos(ph) * std::cos(th)};
    Eigen::Vector3d c;
    for (int i = 0; i < 3; i++) c(i) = y(i) + (1.0 / rho) * m[i];
    return c;
}
double ExtendKF::RandomGenerator(const int low, const int high) { return (std::rand()) / (RAND_MAX + 1.0) * (high - low) + low; }  // :205-219
Eigen::MatrixXd ExtendKF::rand(int row, int column, double min, double max) {  // src/ExtendKF.cpp:220-235
    Eigen::MatrixXd p(row, column);
    for (int i = 0; i < row; i++)
        for (int j = 0; j < column; j++) p(i, j) = (double)std::rand() / RAND_MAX * (max - min) + min;
    return p;
}
void ExtendKF::hinv(Eigen::VectorXd uvd, Eigen::VectorXd Xv, double initial_rho, Eigen::VectorXd& newFeature) {  // src/ExtendKF.cpp:236-265
    const double fku = cam->K(0, 0), fkv = cam->K(1, 1), U0 = cam->K(0, 2), V0 = cam->K(1, 2);
    Eigen::MatrixXd in(2, 1), uv;
    in(0, 0) = uvd(0);
    in(1, 0) = uvd(1);
    undistort_fm(in, uv);
    const double hl[3] = {-(U0 - uv(0, 0)) / fku, -(V0 - uv(1, 0)) / fkv, 1.0};
    Eigen::VectorXd q(4);
    for (int i = 0; i < 4; i++) q(i) = Xv(3 + i);
    Eigen::Matrix3d R = q2r(q);
    double n[3];
    for (int i = 0; i < 3; i++) n[i] = R(i, 0) * hl[0] + R(i, 1) * hl[1] + R(i, 2) * hl[2];
    Eigen::VectorXd out(6);
    for (int i = 0; i < 3; i++) out(i) = Xv(i);
    out(3) = std::atan2(n[0], n[2]);
    out(4) = std::atan2(-n[1], std::sqrt(n[0] * n[0] + n[2] * n[2]));
    out(5) = initial_rho;
    newFeature = out;
}
Eigen::Matrix<double, 3, 4> ExtendKF::dRq_times_a_by_dq(Eigen::VectorXd q, Eigen::Vector3d a) {  // src/ExtendKF.cpp:286-311
    const double q0 = q(0), q1 = q(1), q2 = q(2), q3 = q(3);
    const double M[4][9] = {{q0, -q3, q2, q3, q0, -q1, -q2, q1, q0}, {q1, q2, q3, q2, -q1, -q0, q3, q0, -q1},
                            {-q2, q1, q0, q1, q2, q3, -q0, q3, -q2}, {-q3, -q0, q1, q0, -q3, q2, q1, q2, q3}};
    Eigen::Matrix<double, 3, 4> o;
    for (int k = 0; k < 4; k++)
        for (int i = 0; i < 3; i++) o(i, k) = 2 * (M[k][3 * i] * a(0) + M[k][3 * i + 1] * a(1) + M[k][3 * i + 2] * a(2));
    return o;
}
Eigen::Matrix2d ExtendKF::jacob_undistor_fm(Eigen::VectorXd uvd) {  // src/ExtendKF.cpp:312-332
    const double a = uvd(0) - cam->Cx, b = uvd(1) - cam->Cy, dx = cam->dx, dy = cam->dy;
    const double rd2 = std::pow(a * dx, 2) + std::pow(b * dy, 2);
    const double g = cam->k1 + 2 * cam->k2 * rd2, c = 1 + cam->k1 * rd2 + cam->k2 * rd2 * rd2;
    Eigen::Matrix2d J;
    J(0, 0) = c + a * g * (2 * a * dx * dx);
    J(0, 1) = a * g * (2 * b * dy * dy);
    J(1, 0) = b * g * (2 * a * dx * dx);
    J(1, 1) = c + b * g * (2 * b * dy * dy);
    return J;
}

// ---------------------------------------------------------------------------------------------------------------------
Tracking::Tracking(const std::string&, ExtendKF* m_ExtendKF) : mT_ExtendKF(m_ExtendKF) {}
// the stages of search_IC_matches one by one: the device produces h_i, H_i, S_i (and the warped patches) for all features together
void Tracking::calculate_derivatives(Eigen::VectorXd) {
    rslam_filter* d = mT_ExtendKF->dev_;
    mT_ExtendKF->status_ = d ? rslam_predict_measurements(d) : RSLAM_ERR_INVALID;
}
void Tracking::pred_patch_fc(int, Eigen::Vector3d) {
    rslam_filter* d = mT_ExtendKF->dev_;
    mT_ExtendKF->status_ = d ? rslam_predict_measurements(d) : RSLAM_ERR_INVALID;
}
void Tracking::matching(cv::Mat image) {
    rslam_filter* d = mT_ExtendKF->dev_;
    if (!d) {
        mT_ExtendKF->status_ = RSLAM_ERR_INVALID;
        return;
    }
    int rc = rslam_set_image(d, 0, image.data, image.rows, image.cols, (int)image.step, 0);
    if (rc == 0) rc = rslam_match(d);
    mT_ExtendKF->status_ = rc;
}
void Tracking::calculate_Hi_inverse_depth(Eigen::VectorXd, Eigen::VectorXd, int order, Eigen::MatrixXd& Hi) {
    ExtendKF* kf = mT_ExtendKF;
    if (kf->sync_to_host(false)) return;
    kf->materialize_H(order);
    Hi = kf->features_info[order].H;
}
void Tracking::calculate_Hi_cartesian(Eigen::VectorXd x_v, Eigen::VectorXd yi, int order, Eigen::MatrixXd& Hi) { calculate_Hi_inverse_depth(x_v, yi, order, Hi); }
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

// replay_synth.cpp -- ROS-free restatement of System::TrackRunning's call order (src/System.cpp:111-129) over a synthetic
// sequence dump, driving the drop-in ExtendKF / Map / Tracking classes.  Used by tests/test_gpu_host_classes.py.
//   dump layout (little endian): int32 N, n, T, rows, cols, n_u01 ; double cam9[9] ; double x0[n] ; double P0[n*n] (col-major) ;
//   double patches[N*169] ; then per frame: uint8 image[rows*cols], double u01[n_u01]
//   output: per frame 13 doubles (x_k_k head) + int32 counts {ic, li, hi}
#include <cstdint>
#include <cstdio>
#include <vector>

#include "ransac_slam/ExtendKF.h"
#include "ransac_slam/Map.h"
#include "ransac_slam/Tracking.h"

using namespace ransac_slam;

int main(int argc, char** argv) {
    if (argc < 3) {
        std::fprintf(stderr, "usage: %s dump.bin out.bin\n", argv[0]);
        return 2;
    }
    FILE* in = std::fopen(argv[1], "rb");
    FILE* out = std::fopen(argv[2], "wb");
    if (!in || !out) return 2;
    int32_t hdr[6];
    if (std::fread(hdr, 4, 6, in) != 6) return 2;
    const int N = hdr[0], n = hdr[1], T = hdr[2], rows = hdr[3], cols = hdr[4], n_u01 = hdr[5];
    double cam9[9];
    if (std::fread(cam9, 8, 9, in) != 9) return 2;
    CamParam cam;
    cam.k1 = cam9[0];
    cam.k2 = cam9[1];
    cam.nRows = (int)cam9[2];
    cam.nCols = (int)cam9[3];
    cam.Cx = cam9[4];
    cam.Cy = cam9[5];
    cam.f = cam9[6];
    cam.dx = cam9[7];
    cam.dy = cam9[8];
    ExtendKF kf("", &cam, "constant_velocity");
    Map map(25, &kf);
    map.set_frozen(true);  // the synthetic map is fixed: map_management only resets the per-frame flags
    Tracking tr("", &kf);
    kf.x_k_k.resize(n);
    kf.p_k_k.resize(n, n);
    if (std::fread(kf.x_k_k.data(), 8, n, in) != (size_t)n) return 2;
    if (std::fread(kf.p_k_k.data(), 8, (size_t)n * n, in) != (size_t)n * n) return 2;
    std::vector<double> patches((size_t)N * 169);
    if (std::fread(patches.data(), 8, patches.size(), in) != patches.size()) return 2;
    kf.features_info.resize(N);
    for (int i = 0; i < N; i++) {
        kf.features_info[i].patch_when_matching.resize(13, 13);
        for (int r = 0; r < 13; r++)
            for (int c = 0; c < 13; c++) kf.features_info[i].patch_when_matching(r, c) = patches[(size_t)i * 169 + r * 13 + c];
    }
    if (kf.sync_to_device()) {
        std::fprintf(stderr, "sync_to_device failed: %s\n", rslam_last_error());
        return 1;
    }
    std::vector<uint8_t> img((size_t)rows * cols);
    std::vector<double> u(n_u01);
    for (int t = 0; t < T; t++) {
        if (std::fread(img.data(), 1, img.size(), in) != img.size()) return 2;
        if (std::fread(u.data(), 8, n_u01, in) != (size_t)n_u01) return 2;
        cv::Mat image(rows, cols, img.data(), (size_t)cols);
        // --- System::TrackRunning (src/System.cpp:111-129) ---
        map.map_management(image, t + 1);
        kf.ekf_prediction();
        tr.search_IC_matches(image);
        tr.set_uniform_draws(u.data(), n_u01);
        tr.ransac_hypotheses();
        kf.ekf_update_li_inliers();
        tr.rescue_hi_inliers();
        kf.ekf_update_hi_inliers();
        if (kf.last_status()) {
            std::fprintf(stderr, "frame %d failed: %s\n", t, rslam_last_error());
            return 1;
        }
        kf.sync_to_host(false);
        int32_t cnt[3] = {0, 0, 0};
        for (const Feature& f : kf.features_info) {
            cnt[0] += f.individually_compatible;
            cnt[1] += f.low_innovation_inlier;
            cnt[2] += f.high_innovation_inlier;
        }
        std::fwrite(kf.x_k_k.data(), 8, 13, out);
        std::fwrite(cnt, 4, 3, out);
    }
    std::fclose(in);
    std::fclose(out);
    return 0;
}

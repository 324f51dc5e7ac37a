// replay_pgm.cpp -- ROS-free replay of an image sequence through the drop-in ExtendKF / Map / Tracking classes in exactly
// System::System's construction order (src/System.cpp:22-70) and System::TrackRunning's call order (src/System.cpp:111-129); the
// main loop of examples/Monocular/mono_slam.cpp:47-69 with cv::imread replaced by a P5 (binary PGM) reader.  Configuration C1.
//   usage: rslam_replay_pgm settings.yaml out.bin [--draws draws.bin] frame0.pgm frame1.pgm ...
//   draws.bin (optional): per frame 100 + 1000 little-endian int32 libc-style draws r in [0, RAND_MAX): the first 100 feed
//     Map::initialize_features (2 per attempt), the next 1000 Tracking::ransac_hypotheses, each as u = r / RAND_MAX like
//     ExtendKF::rand (src/ExtendKF.cpp:230).  Without it the classes draw from std::rand() like the reference.
//   out.bin: per frame 13 doubles (camera state) + int32 {N features, ic, li, hi, deleted, initialised}
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "ransac_slam/ExtendKF.h"
#include "ransac_slam/Map.h"
#include "ransac_slam/Tracking.h"

using namespace ransac_slam;

static bool read_pgm(const char* path, std::vector<uint8_t>& px, int* rows, int* cols) {
    FILE* f = std::fopen(path, "rb");
    if (!f) return false;
    char magic[3] = {0, 0, 0};
    int vals[3], got = 0;
    if (std::fscanf(f, "%2s", magic) != 1 || std::strcmp(magic, "P5") != 0) {
        std::fclose(f);
        return false;
    }
    while (got < 3) {  // width, height, maxval, with '#' comment lines in between
        int c = std::fgetc(f);
        if (c == EOF) break;
        if (c == '#') {
            while (c != '\n' && c != EOF) c = std::fgetc(f);
        } else if (c >= '0' && c <= '9') {
            std::ungetc(c, f);
            if (std::fscanf(f, "%d", &vals[got]) != 1) break;
            got++;
        }
    }
    if (got != 3 || vals[2] != 255) {
        std::fclose(f);
        return false;
    }
    std::fgetc(f);  // the single whitespace byte after maxval
    *cols = vals[0];
    *rows = vals[1];
    px.resize((size_t)vals[0] * vals[1]);
    const bool ok = std::fread(px.data(), 1, px.size(), f) == px.size();
    std::fclose(f);
    return ok;
}

int main(int argc, char** argv) {
    if (argc < 4) {
        std::fprintf(stderr, "usage: %s settings.yaml out.bin [--draws draws.bin] frame0.pgm ...\n", argv[0]);
        return 2;
    }
    const std::string yaml = argv[1];
    FILE* out = std::fopen(argv[2], "wb");
    FILE* draws = nullptr;
    int first = 3;
    if (std::strcmp(argv[3], "--draws") == 0 && argc > 5) {
        draws = std::fopen(argv[4], "rb");
        if (!draws) return 2;
        first = 5;
    }
    if (!out) return 2;
    // --- System::System (src/System.cpp:22-70) ---
    CamParam cam;
    int min_features = 25;
    if (!load_camera_yaml(yaml, &cam, &min_features)) {
        std::fprintf(stderr, "Failed to open settings file at: %s\n", yaml.c_str());
        return 1;
    }
    ExtendKF kf(yaml, &cam, "constant_velocity");
    kf.initialize_x_and_p();
    Map map(min_features, &kf);
    Tracking tr(yaml, &kf);
    long long TrackSteps = 0;
    std::vector<uint8_t> px;
    std::vector<int32_t> r(1100);
    std::vector<double> u(1100);
    for (int a = first; a < argc; a++) {
        int rows = 0, cols = 0;
        if (!read_pgm(argv[a], px, &rows, &cols)) {
            std::fprintf(stderr, "Failed to load image at: %s\n", argv[a]);
            return 1;
        }
        cv::Mat image(rows, cols, px.data(), (size_t)cols);
        if (draws) {
            if (std::fread(r.data(), 4, r.size(), draws) != r.size()) return 2;
            for (size_t i = 0; i < r.size(); i++) u[i] = (double)r[i] / (double)RAND_MAX;
            map.set_uniform_draws(u.data(), 100);
            tr.set_uniform_draws(u.data() + 100, 1000);
        }
        // --- System::TrackRunning (src/System.cpp:103-129) ---
        TrackSteps++;
        map.map_management(image, (int)TrackSteps);
        kf.ekf_prediction();
        tr.search_IC_matches(image);
        tr.ransac_hypotheses();
        kf.ekf_update_li_inliers();
        tr.rescue_hi_inliers();
        kf.ekf_update_hi_inliers();
        if (kf.last_status() < 0) {
            std::fprintf(stderr, "frame %lld failed: %s\n", TrackSteps, rslam_last_error());
            return 1;
        }
        kf.sync_to_host(false);
        int32_t cnt[6] = {(int32_t)kf.features_info.size(), 0, 0, 0, map.last_info()[0], map.last_info()[2]};
        for (const Feature& f : kf.features_info) {
            cnt[1] += f.individually_compatible;
            cnt[2] += f.low_innovation_inlier;
            cnt[3] += f.high_innovation_inlier;
        }
        std::fwrite(kf.x_k_k.data(), 8, 13, out);
        std::fwrite(cnt, 4, 6, out);
    }
    std::fclose(out);
    if (draws) std::fclose(draws);
    return 0;
}

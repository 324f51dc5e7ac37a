"""The host classes drop in for the reference's: (1) the reference's OWN src/System.cpp and src/Converter.cpp compile against
ransac_slam_b200/host/ransac_slam/{ExtendKF,Map,Tracking}.h (the reference's System.h included through the dispatcher header) and
link with host_classes.cpp + librslam_b200.so -- the recipe of INTEGRATION.md, with the stand-in Eigen / OpenCV / ROS headers of
oracle/ref_shim in place of the real ones; (2) the public camera-model helpers of the host ExtendKF
(include/ransac_slam/ExtendKF.h:57-143) return what the formulas of the reference give (numpy restatement in ransac_slam_b200/synth.py).
CPU only: nothing here needs a GPU (no filter is stepped)."""
import os
import subprocess

import numpy as np
import pytest

from ransac_slam_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
HOST = os.path.join(ROOT, "ransac_slam_b200", "host")
LIB = os.path.join(ROOT, "ransac_slam_b200", "lib")
SHIM = os.path.join(ROOT, "oracle", "ref_shim")
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _flags():
    return ["-std=c++17", "-O1", "-w", "-fPIC", "-I", HOST, "-I", SHIM, "-I", os.path.join(REF, "include"),
            '-DRSLAM_REFERENCE_SYSTEM_H="%s"' % os.path.join(REF, "include", "ransac_slam", "System.h")]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="the reference tree is not on this machine")
def test_reference_system_cpp_links_against_the_host_classes(tmp_path):
    objs = []
    srcs = [os.path.join(REF, "src", "System.cpp"), os.path.join(REF, "src", "Converter.cpp"), os.path.join(HOST, "host_classes.cpp")]
    main = tmp_path / "main.cpp"
    main.write_text('#include "ransac_slam/System.h"\n'
                    "int main(int argc, char** argv) {\n"
                    "    if (argc < 2) return 0;  // link check: System::System opens ROS topics and waits for a subscriber\n"
                    "    ransac_slam::System SLAM(argv[1], true);\n"
                    "    cv::Mat im;\n"
                    "    SLAM.TrackRunning(im);\n"
                    "    return 0;\n}\n")
    for src in srcs + [str(main)]:
        obj = str(tmp_path / (os.path.basename(src) + ".o"))
        r = subprocess.run([GXX] + _flags() + ["-c", src, "-o", obj], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
        objs.append(obj)
    exe = str(tmp_path / "mono_dropin")
    r = subprocess.run([GXX, "-o", exe] + objs + ["-L", LIB, "-lrslam_b200", "-Wl,-rpath," + LIB, "-Wl,--no-undefined"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    # every method System::TrackRunning calls (src/System.cpp:111-129) resolved to the GPU-backed classes
    nm = subprocess.run(["nm", "-C", exe], capture_output=True, text=True).stdout
    for sym in ("ransac_slam::Map::map_management", "ransac_slam::ExtendKF::ekf_prediction", "ransac_slam::Tracking::search_IC_matches",
                "ransac_slam::Tracking::ransac_hypotheses", "ransac_slam::ExtendKF::ekf_update_li_inliers", "ransac_slam::Tracking::rescue_hi_inliers",
                "ransac_slam::ExtendKF::ekf_update_hi_inliers", "ransac_slam::ExtendKF::inversedepth2cartesian", "rslam_map_management", "rslam_update_hi"):
        assert sym in nm, sym
    assert subprocess.run([exe]).returncode == 0


_HELPERS = r'''
#include <cstdio>
#include "ransac_slam/ExtendKF.h"
using namespace ransac_slam;
static void pr(const char* n, const double* v, int k) { printf("%s", n); for (int i = 0; i < k; i++) printf(" %.17g", v[i]); printf("\n"); }
int main() {
    CamParam cam;
    cam.k1 = 0.06333; cam.k2 = 0.01390; cam.nRows = 240; cam.nCols = 320; cam.dx = cam.dy = 0.0112; cam.f = 2.1735;
    cam.Cx = 1.7945 / 0.0112; cam.Cy = 1.4433 / 0.0112;
    const double K[9] = {cam.f / cam.dx, 0, cam.Cx, 0, cam.f / cam.dy, cam.Cy, 0, 0, 1};
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) cam.K(i, j) = K[3 * i + j];
    ExtendKF kf("", &cam, "constant_velocity");
    Eigen::VectorXd q(4); q(0) = 0.9; q(1) = 0.1; q(2) = -0.3; q(3) = 0.2;
    Eigen::Matrix3d R = kf.q2r(q);
    double r9[9]; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r9[3 * i + j] = R(i, j);
    pr("q2r", r9, 9);
    Eigen::MatrixXd uv(2, 3), o;
    uv(0, 0) = 12.5; uv(1, 0) = 200.25; uv(0, 1) = 160; uv(1, 1) = 128; uv(0, 2) = 310.75; uv(1, 2) = 7.5;
    kf.distort_fm(uv, o);
    double d6[6] = {o(0, 0), o(1, 0), o(0, 1), o(1, 1), o(0, 2), o(1, 2)};
    pr("distort", d6, 6);
    kf.undistort_fm(uv, o);
    double u6[6] = {o(0, 0), o(1, 0), o(0, 1), o(1, 1), o(0, 2), o(1, 2)};
    pr("undistort", u6, 6);
    Eigen::VectorXd y(6); y(0) = 0.1; y(1) = -0.2; y(2) = 0.3; y(3) = 0.25; y(4) = -0.15; y(5) = 0.2;
    Eigen::Vector3d c = kf.inversedepth2cartesian(y);
    double c3[3] = {c(0), c(1), c(2)};
    pr("id2c", c3, 3);
    Eigen::Vector3d p; p(0) = 0.3; p(1) = -0.2; p(2) = 2.0;
    Eigen::Vector2d h = kf.hu(p);
    double h2[2] = {h(0), h(1)};
    pr("hu", h2, 2);
    Eigen::VectorXd uvd(2), Xv(13), nf; uvd(0) = 201; uvd(1) = 77;
    for (int i = 0; i < 13; i++) Xv(i) = 0.01 * (i + 1);
    Xv(3) = 0.9; Xv(4) = 0.1; Xv(5) = -0.3; Xv(6) = 0.2;
    kf.hinv(uvd, Xv, 1.0, nf);
    double n6[6]; for (int i = 0; i < 6; i++) n6[i] = nf(i);
    pr("hinv", n6, 6);
    Eigen::Matrix2d J = kf.jacob_undistor_fm(uvd);
    double j4[4] = {J(0, 0), J(0, 1), J(1, 0), J(1, 1)};
    pr("jac", j4, 4);
    Eigen::Matrix<double, 3, 4> D = kf.dRq_times_a_by_dq(q, p);
    double d12[12]; for (int i = 0; i < 3; i++) for (int j = 0; j < 4; j++) d12[4 * i + j] = D(i, j);
    pr("dRq", d12, 12);
    Eigen::MatrixXd zi;
    kf.hi_cartesian(p, zi);
    double z2[2] = {zi(0, 0), zi(1, 0)};
    pr("hi", z2, 2);
    Eigen::Vector3d behind; behind(0) = 0.1; behind(1) = 0.1; behind(2) = -1.0;
    kf.hi_cartesian(behind, zi);
    printf("hi_behind %d\n", (int)(zi.rows() * zi.cols()));
    return 0;
}
'''


@pytest.mark.parametrize("mode", ["standalone", "package"])
def test_host_helpers_follow_the_reference_formulas(tmp_path, mode):
    """both header modes: the repository's own minimal linear-algebra shim, and (where the reference tree exists) the reference's System.h
    over the stand-in Eigen"""
    if mode == "package" and not os.path.isdir(os.path.join(REF, "include")):
        pytest.skip("the reference tree is not on this machine")
    src = tmp_path / "helpers.cpp"
    src.write_text(_HELPERS)
    exe = str(tmp_path / "helpers")
    flags = _flags() if mode == "package" else ["-std=c++17", "-O1", "-w", "-I", HOST]
    r = subprocess.run([GXX] + flags + [str(src), os.path.join(HOST, "host_classes.cpp"), "-o", exe, "-L", LIB, "-lrslam_b200", "-Wl,-rpath," + LIB],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    got = {ln.split()[0]: np.array([float(v) for v in ln.split()[1:]]) for ln in out.stdout.splitlines()}
    cam = synth.Camera()
    q = np.array([0.9, 0.1, -0.3, 0.2])
    np.testing.assert_allclose(got["q2r"].reshape(3, 3), synth.q2r(q), rtol=1e-15)
    uv = np.array([[12.5, 200.25], [160.0, 128.0], [310.75, 7.5]])
    np.testing.assert_allclose(got["distort"].reshape(3, 2), synth.distort(cam, uv), rtol=1e-13)
    np.testing.assert_allclose(got["undistort"].reshape(3, 2), synth.undistort(cam, uv), rtol=1e-13)
    y = np.array([0.1, -0.2, 0.3, 0.25, -0.15, 0.2])
    m = np.array([np.cos(y[4]) * np.sin(y[3]), -np.sin(y[4]), np.cos(y[4]) * np.cos(y[3])])
    np.testing.assert_allclose(got["id2c"], y[:3] + m / y[5], rtol=1e-14)
    p = np.array([0.3, -0.2, 2.0])
    np.testing.assert_allclose(got["hu"], [cam.Cx + p[0] / p[2] * cam.fku, cam.Cy + p[1] / p[2] * cam.fkv], rtol=1e-14)
    Xv = 0.01 * np.arange(1, 14)
    Xv[3:7] = q
    np.testing.assert_allclose(got["hinv"], synth.hinv(cam, np.array([201.0, 77.0]), Xv, 1.0), rtol=1e-13)
    np.testing.assert_allclose(got["jac"].reshape(2, 2), synth.jacob_undistort(cam, np.array([201.0, 77.0])), rtol=1e-13)
    np.testing.assert_allclose(got["dRq"].reshape(3, 4), synth.dRq_times_a_by_dq(q, p), rtol=1e-14)
    hd = synth.distort(cam, np.array([[cam.Cx + p[0] / p[2] * cam.fku, cam.Cy + p[1] / p[2] * cam.fkv]]))[0]
    np.testing.assert_allclose(got["hi"], hd, rtol=1e-13)
    assert got["hi_behind"][0] == 0  # behind the camera: the +-60 degree gate empties zi (src/ExtendKF.cpp:106-113)

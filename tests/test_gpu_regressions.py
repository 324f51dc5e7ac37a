"""Regression tests for defects found in review, GPU vs oracle through the C ABI:
  * erasing a feature record moves ALL of its members (per-frame flags included): counters and the `measured` count of the next
    Map::map_management (src/Map.cpp:34-66) must follow the oracle after deletions with mixed flags;
  * Tracking::rescue_hi_inliers re-linearises, at x_k_k, a matched feature that has left the field of view since x_k_km1, with its
    stale h as the linearisation pixel (src/ExtendKF.cpp:77-78, src/Tracking.cpp:553-565, 578-579);
  * the fixed-point Newton distortion of the prediction kernel against the reference's ten IEEE steps: last-bit agreement of h."""
import numpy as np
import pytest

from oracle import oracle_py as O
from ransac_slam_b200 import synth
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _blank(scene, seq, k, feats):
    """paint the background over the pasted appearance of some features in frame k: they are predicted but not matched"""
    bg = synth.background(scene.cam)
    for i in feats:
        x, y = seq.z_true[k][i]
        if x >= 0:
            seq.images[k][y - 6:y + 7, x - 6:x + 7] = bg[y - 6:y + 7, x - 6:x + 7]


def test_deletion_moves_flags_with_the_record():
    scene, x, P = synth.random_spd_state(30, seed=33)
    seq = synth.make_sequence(scene, T=3, seed=35, t0=3)
    _blank(scene, seq, 0, [3, 7, 11, 19, 22, 27])  # six features without a match in the first frame
    o = H.oracle_from(scene, x, P, prior=False, sparse=False)
    g = H.gpu_from(scene, x, P, prior=False, max_features=34)
    o.frame(seq.images[0], seq.u01[0])
    g.frame(seq.images[0][None], seq.u01[0][None])
    fo, fg = o.features(), g.features()
    for k in ("has_h", "ic", "li", "hi"):
        assert (fo[k] == fg[k]).all()
    flags = fo["li"] | fo["hi"]
    assert 3 < flags.sum() < 30 and fo["has_h"].sum() > flags.sum()  # mixed flags, or the test says nothing
    # make three features deletable: one measured, one predicted-only, one in between the survivors with mixed flags
    meas = np.flatnonzero(flags)
    unm = np.flatnonzero(~flags)
    kill = sorted([int(meas[1]), int(unm[0]), int(meas[-1])])
    tp = np.full(30, 10, np.int32)
    tm = np.full(30, 9, np.int32)
    tm[kill] = 1
    o.set_counters(tp, tm)
    g.set_counters(tp, tm)
    # Map::map_management steps 1 + 2: delete pass, then counters from the flags that moved with the surviving records
    rco, ndo = o.map_delete_pass(False)
    rcg, ndg = g.map_delete_features(False)
    assert (rco, ndo) == (rcg, ndg) == (0, 3)
    fo, fg = o.features(), g.features()
    for k in ("has_h", "ic", "li", "hi"):
        assert (fo[k] == fg[k]).all(), k
    assert np.array_equal(fo["z"][fo["ic"]], fg["z"][fg["ic"]])
    np.testing.assert_allclose(fg["h"][fg["has_h"]], fo["h"][fo["has_h"]], rtol=0, atol=1e-9)
    o.map_reset_flags()
    g.begin_frame()
    fo, fg = o.features(), g.features()
    assert list(fo["times_predicted"]) == list(fg["times_predicted"]) and list(fo["times_measured"]) == list(fg["times_measured"])
    assert fo["times_measured"].sum() == 9 * 27 + int(np.delete(flags, kill).sum())
    # and the following frames keep agreeing (state, flags, counters)
    for k in (1, 2):
        o.ekf_prediction(); g.ekf_prediction()
        o.search_ic_matches(seq.images[k]); g.set_image(seq.images[k]); g.search_ic_matches()
        o.ransac_hypotheses(seq.u01[k]); g.ransac_hypotheses(seq.u01[k])
        o.update_li(); g.update_li()
        o.rescue_hi(); g.rescue_hi()
        o.update_hi(); g.update_hi()
        o.map_reset_flags(); g.begin_frame()
        fo, fg = o.features(), g.features()
        assert list(fo["times_predicted"]) == list(fg["times_predicted"]) and list(fo["times_measured"]) == list(fg["times_measured"]), k
    xo, Po = o.get_state()
    xg, Pg = g.download_state()
    H.assert_x_close(xg, xo)
    H.assert_P_close(Pg, Po)


def test_map_management_measured_count_after_deletion():
    """the whole Map::map_management: `measured` (src/Map.cpp:57-66) decides how many features are initialised"""
    scene, x, P = synth.random_spd_state(20, seed=43)
    seq = synth.make_sequence(scene, T=2, seed=45, t0=3)
    _blank(scene, seq, 0, [2, 9, 15])
    o = H.oracle_from(scene, x, P, prior=False, sparse=False)
    g = H.gpu_from(scene, x, P, prior=False, max_features=60)
    o.frame(seq.images[0], seq.u01[0])
    g.frame(seq.images[0][None], seq.u01[0][None])
    flags = o.features()["li"] | o.features()["hi"]
    tp = np.full(20, 10, np.int32)
    tm = np.full(20, 9, np.int32)
    tm[[int(np.flatnonzero(flags)[0]), int(np.flatnonzero(~flags)[-1])]] = 0
    o.set_counters(tp, tm)
    g.set_counters(tp, tm)
    u = np.random.default_rng(3).random(100)
    min_features = int(flags.sum()) + 2  # measured survivors < min_features -> initialise (min_features - measured) features
    rco, io = o.map_management(seq.images[1], 2, min_features, u, reference_indexing=False)
    g.set_image(seq.images[1])
    rcg, ig = g.map_management(2, min_features, u, reference_indexing=False)
    assert rco == rcg == 0
    assert io == ig, (io, ig)
    assert ig["deleted"] == 2
    fo, fg = o.features(), g.features()
    assert list(fo["times_predicted"]) == list(fg["times_predicted"]) and list(fo["times_measured"]) == list(fg["times_measured"])
    assert list(o.types()) == list(g.types())


def test_rescue_relinearises_a_feature_that_left_the_view():
    scene, x, P = synth.random_spd_state(16, seed=51)
    cam = scene.cam
    # move feature 3's ray next to the right image border at x_k_km1 so that a small yaw of x_k_k pushes it outside
    o = H.oracle_from(scene, x, P, sparse=False)
    o.search_ic_matches(None)
    # yaw until the feature with the largest u is < 0.5 px inside the border
    ho = o.features()["h"]
    j = int(np.argmax(ho[:, 0]))
    x1 = x.copy()
    lo, hi_ = 0.0, 0.5
    for _ in range(60):  # bisection on the yaw angle about +y that leaves feature j just inside the right border
        a = 0.5 * (lo + hi_)
        q = np.array([np.cos(a / 2), 0.0, -np.sin(a / 2), 0.0])
        xt = x.copy()
        xt[3:7] = _qmul(x[3:7], q)
        h, vis = _predict(cam, xt)
        if vis[j] and h[j, 0] < cam.nCols - 0.4:
            lo = a
        else:
            hi_ = a
    q = np.array([np.cos(lo / 2), 0.0, -np.sin(lo / 2), 0.0])
    x1[3:7] = _qmul(x[3:7], q)
    h1, vis1 = _predict(cam, x1)
    assert vis1[j] and h1[j, 0] > cam.nCols - 1.0
    # posterior pose: a little more yaw -> feature j fails the image gate at x_k_k
    q2 = np.array([np.cos(0.01), 0.0, -np.sin(0.01), 0.0])
    x2 = x1.copy()
    x2[3:7] = _qmul(x1[3:7], q2)
    h2, vis2 = _predict(cam, x2)
    assert not vis2[j] and vis2.sum() >= 8
    o = H.oracle_from(scene, x1, P, sparse=False)
    g = H.gpu_from(scene, x1, P)
    o.search_ic_matches(None)
    g.search_ic_matches()
    fo = o.features()
    assert fo["has_h"][j]
    z = np.rint(fo["h"]) + np.array([1.0, -1.0])
    ic = fo["has_h"].astype(np.uint8)
    o.set_matches(z, ic)
    g.set_matches(z, ic)
    # no low-innovation inliers: every match goes through the rescue gate at x_k_k = x2
    o.set_state(x2, P, prior=False)
    g.upload_state(x2, P, prior=False)
    ho_stale, hg_stale = fo["h"][j].copy(), g.features()["h"][j].copy()
    o.rescue_hi()
    g.rescue_hi()
    fo, fg = o.features(), g.features()
    assert fg["has_h"][j] and fo["has_h"][j] and np.array_equal(fg["h"][j], hg_stale) and np.array_equal(fo["h"][j], ho_stale)  # h of x_k_km1 stays
    others = np.flatnonzero(fo["has_h"])
    others = others[others != j]
    assert not np.array_equal(fg["h"][others], np.rint(fg["h"][others]))  # ... while the visible ones were re-predicted
    assert (fo["hi"] == fg["hi"]).all()
    Hc, Hf = g.H_sparse()
    for i in np.flatnonzero(fo["has_h"]):
        Hd = o.H_dense(int(i))
        np.testing.assert_allclose(Hc[i], Hd[:, :7], rtol=1e-9, atol=1e-11, err_msg=f"feature {i}")
        np.testing.assert_allclose(Hf[i], Hd[:, 13 + 6 * i:19 + 6 * i], rtol=1e-9, atol=1e-11, err_msg=f"feature {i}")
    o.update_hi()
    g.update_hi()
    xo, Po = o.get_state()
    xg, Pg = g.download_state()
    H.assert_x_close(xg, xo)
    H.assert_P_close(Pg, Po)


def _qmul(a, b):
    r1, x1, y1, z1 = a
    r2, x2, y2, z2 = b
    return np.array([r1 * r2 - x1 * x2 - y1 * y2 - z1 * z2, r1 * x2 + x1 * r2 + y1 * z2 - z1 * y2, r1 * y2 - x1 * z2 + y1 * r2 + z1 * x2,
                     r1 * z2 + x1 * y2 - y1 * x2 + z1 * r2])


def _predict(cam, x):
    from oracle import np_oracle as NP

    N = (x.size - 13) // 6
    return NP.predict_h(cam.as9(), x, np.zeros(N, int))


def test_fixed_point_distortion_against_ten_ieee_steps():
    """k_predict's Newton iteration stops at its fixed point and uses a reciprocal-based step; the reference runs ten IEEE-division
    steps (src/ExtendKF.cpp:191-196).  h must agree to the last bits: discrete decisions downstream (round(h), gates) then differ only
    at rounding boundaries."""
    worst = 0.0
    for seed in (61, 62, 63):
        scene, x, P = synth.random_spd_state(100, seed=seed)
        o = H.oracle_from(scene, x, P, sparse=True)
        g = H.gpu_from(scene, x, P)
        o.search_ic_matches(None)
        g.search_ic_matches()
        fo, fg = o.features(), g.features()
        assert (fo["has_h"] == fg["has_h"]).all()
        v = fo["has_h"]
        ulp = np.abs(fg["h"][v] - fo["h"][v]) / np.spacing(np.abs(fo["h"][v]))
        worst = max(worst, float(ulp.max()))
        assert (np.rint(fg["h"][v]) == np.rint(fo["h"][v])).all()
    print(f"max |h_gpu - h_oracle| = {worst:.1f} ulp")
    assert worst <= 16.0

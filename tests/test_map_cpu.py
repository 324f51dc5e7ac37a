"""Oracle restatement of the reference's Map state surgery (src/Map.cpp:19-32, 69-196, 268-311, 339-400) against an independent
numpy restatement and the properties the operations must have.  CPU only."""
import numpy as np

from oracle import oracle_py as O
from ransac_slam_b200 import synth
from tests import helpers as H


def _offsets(types):
    off, out = 13, []
    for t in types:
        out.append(off)
        off += 6 if t == 0 else 3
    return out, off


def test_delete_feature_is_row_column_removal():
    scene, x, P = synth.random_spd_state(12, seed=3)
    o = H.oracle_from(scene, x, P, prior=False)
    assert o.map_delete_feature(4) == 0
    xo, Po = o.get_state()
    keep = np.r_[0:13 + 6 * 4, 13 + 6 * 5:x.size]
    assert np.array_equal(xo, x[keep]) and np.array_equal(Po, P[np.ix_(keep, keep)])
    assert o.N == 11 and o.n == x.size - 6


def test_delete_pass_reference_indexing_runs_ahead_after_first_erase():
    """src/Map.cpp:22-31: `i` advances on every iteration, the iterator only when nothing was erased.  With features 2 and 5 (0-based)
    failing the test, the reference erases records 2 and 5 but removes the STATE blocks of features 2 and 6."""
    scene, x, P = synth.random_spd_state(10, seed=5)
    tp = np.full(10, 10, np.int32)
    tm = np.full(10, 9, np.int32)
    tm[[2, 5]] = 1
    o = H.oracle_from(scene, x, P, prior=False)
    o.set_counters(tp, tm)
    rc, nd = o.map_delete_pass(reference_indexing=True)
    assert rc == 0 and nd == 2 and o.N == 8
    xo, Po = o.get_state()
    blk = lambda i: np.r_[13 + 6 * i:13 + 6 * i + 6]
    keep = np.setdiff1d(np.arange(x.size), np.r_[blk(2), blk(6)])
    assert np.array_equal(xo, x[keep]) and np.array_equal(Po, P[np.ix_(keep, keep)])
    # consistent variant removes what it erased
    o2 = H.oracle_from(scene, x, P, prior=False)
    o2.set_counters(tp, tm)
    rc, nd = o2.map_delete_pass(reference_indexing=False)
    keep2 = np.setdiff1d(np.arange(x.size), np.r_[blk(2), blk(5)])
    x2, P2 = o2.get_state()
    assert rc == 0 and nd == 2 and np.array_equal(x2, x[keep2]) and np.array_equal(P2, P[np.ix_(keep2, keep2)])


def test_delete_pass_last_feature_is_reference_ub():
    scene, x, P = synth.random_spd_state(6, seed=7)
    tp = np.full(6, 10, np.int32)
    tm = np.full(6, 9, np.int32)
    tm[5] = 0
    o = H.oracle_from(scene, x, P, prior=False)
    o.set_counters(tp, tm)
    rc, _ = o.map_delete_pass(reference_indexing=True)
    assert rc == -4  # delete_a_feature(6) reads features_info[5] of a 5-element vector (src/Map.cpp:73)
    o2 = H.oracle_from(scene, x, P, prior=False)
    o2.set_counters(tp, tm)
    rc, nd = o2.map_delete_pass(reference_indexing=False)
    assert rc == 0 and nd == 1 and o2.n == x.size - 6


def _np_convert(x, P, types, thr=0.1):
    offs, n = _offsets(types)
    for i, t in enumerate(types):
        if t != 0:
            continue
        ip = offs[i]
        rho, th, ph = x[ip + 5], x[ip + 3], x[ip + 4]
        m = np.array([np.cos(ph) * np.sin(th), -np.sin(ph), np.cos(ph) * np.cos(th)])
        xo = x[ip:ip + 3] + m / rho
        std_d = np.sqrt(P[ip + 5, ip + 5]) / rho**2
        d1, d2 = xo - x[ip:ip + 3], xo - x[0:3]
        li = 4 * std_d * (d1 @ d2) / (np.linalg.norm(d1) * np.linalg.norm(d2)) / np.linalg.norm(d2)
        if li < thr:
            J = np.zeros((3, 6))
            J[:, :3] = np.eye(3)
            J[:, 3] = np.array([np.cos(ph) * np.cos(th), 0, -np.cos(ph) * np.sin(th)]) / rho
            J[:, 4] = np.array([-np.sin(ph) * np.sin(th), -np.cos(ph), -np.sin(ph) * np.cos(th)]) / rho
            J[:, 5] = -m / rho**2
            T = np.zeros((n - 3, n))
            T[:ip, :ip] = np.eye(ip)
            T[ip:ip + 3, ip:ip + 6] = J
            T[ip + 3:, ip + 6:] = np.eye(n - ip - 6)
            xn = np.r_[x[:ip], xo, x[ip + 6:]]
            return i, xn, T @ P @ T.T
    return -1, x, P


def test_inversedepth_to_cartesian_matches_numpy():
    scene, x, P = synth.random_spd_state(10, seed=11)
    # make features 3 and 7 well localised (tiny rho variance -> small linearity index); the reference converts only the first
    for i in (3, 7):
        ip = 13 + 6 * i
        s = 1e-3 / np.sqrt(P[ip + 5, ip + 5])
        P[ip + 5, :] *= s
        P[:, ip + 5] *= s
    o = H.oracle_from(scene, x, P, prior=False)
    idx = o.map_inversedepth_to_cartesian()
    i_np, x_np, P_np = _np_convert(x, P, [0] * 10)
    assert idx == i_np == 3
    xo, Po = o.get_state()
    np.testing.assert_allclose(xo, x_np, rtol=1e-13, atol=0)
    np.testing.assert_allclose(Po, P_np, rtol=1e-12, atol=1e-18)
    assert list(o.types()) == [0, 0, 0, 1, 0, 0, 0, 0, 0, 0]
    # second call converts the next candidate; a third finds none
    assert o.map_inversedepth_to_cartesian() == 7
    assert o.map_inversedepth_to_cartesian() == -1


def test_add_feature_matches_python_restatement_and_reprojects():
    scene, x, P = synth.random_spd_state(8, seed=13)
    o = H.oracle_from(scene, x, P, prior=False)
    uv = np.array([201.0, 77.0])
    img = synth.background(scene.cam, seed=3)
    idx = o.map_add_feature(uv, img)
    assert idx == 8 and o.N == 9 and o.n == x.size + 6
    xo, Po = o.get_state()
    y = synth.hinv(scene.cam, uv, x[:13], 1.0)
    np.testing.assert_allclose(xo[-6:], y, rtol=1e-14)
    dy_dxv, dy_dhd = synth.feature_init_jacobians(scene.cam, uv, x[:13], reference_fill=True)
    n = x.size
    Padd = np.diag([scene.std_z**2, scene.std_z**2, 1.0])
    np.testing.assert_array_equal(Po[:n, :n], P)
    np.testing.assert_allclose(Po[n:, :n], dy_dxv @ P[:13, :n], rtol=1e-12, atol=1e-18)
    np.testing.assert_allclose(Po[:n, n:], (dy_dxv @ P[:13, :n]).T, rtol=1e-12, atol=1e-18)
    np.testing.assert_allclose(Po[n:, n:], dy_dxv @ P[:13, :13] @ dy_dxv.T + dy_dhd @ Padd @ dy_dhd.T, rtol=1e-12, atol=1e-18)
    # the new feature projects back onto the pixel it was initialised from
    m = np.array([np.cos(y[4]) * np.sin(y[3]), -np.sin(y[4]), np.cos(y[4]) * np.cos(y[3])])
    uvp, _ = synth.project(scene.cam, x[0:3], x[3:7], (y[0:3] + m / y[5])[None])
    np.testing.assert_allclose(uvp[0], uv, atol=1e-6)
    patch, pose = o.feature_init(8)
    assert np.array_equal(patch, img[77 - 20:77 + 21, 201 - 20:201 + 21])
    np.testing.assert_allclose(pose[:3], x[:3])
    np.testing.assert_allclose(pose[3:12].reshape(3, 3), synth.q2r(x[3:7]))
    assert tuple(pose[12:]) == (201.0, 77.0)


def test_fast9_matches_cv2_fixtures():
    """the oracle's cv::FAST restatement against keypoints produced by cv2 4.13 (tests/golden/make_fast_fixtures.py), order included"""
    import os

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fast_fixtures.npz"))
    total = 0
    for name in g["names"]:
        for t in (100, 40):
            kp = O.fast9(g[f"{name}_img"], t, True, 20000)
            ref = g[f"{name}_kp{t}"]
            assert kp.shape == ref.shape and (kp == ref).all(), (name, t)
            total += len(ref)
    assert total > 300
    p = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pgm_frames.npz"))
    for i in range(p["frames"].shape[0]):  # frames of the reference's bundled sequence
        kp = O.fast9(p["frames"][i], 100, True, 20000)
        assert kp.shape == p[f"kp{i}"].shape and (kp == p[f"kp{i}"]).all()


def test_map_management_bootstraps_a_map_from_nothing():
    """first frame: no features -> measured == 0 -> initialize_features(min_features) (src/Map.cpp:58-62)"""
    import os

    p = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pgm_frames.npz"))
    cam = synth.Camera()
    o = O.OracleFilter(cam.as9())
    o.initialize_x_and_p()
    u = np.random.default_rng(5).random(200)
    rc, info = o.map_management(p["frames"][0], 1, 25, u)
    assert rc == 0 and info["deleted"] == 0 and info["converted"] == -1
    assert 0 < info["initialised"] <= 25 and info["attempts"] <= 50 and o.N == info["initialised"] and o.n == 13 + 6 * o.N
    x, P = o.get_state()
    assert np.allclose(P, P.T) and np.all(np.diag(P)[13:] >= 0)
    for i in range(o.N):  # every new feature sits where FAST found a corner, shifted by the port's (-1, -1)
        patch, pose = o.feature_init(i)
        uv = pose[12:].astype(int)
        assert np.array_equal(patch, p["frames"][0][uv[1] - 20:uv[1] + 21, uv[0] - 20:uv[0] + 21])

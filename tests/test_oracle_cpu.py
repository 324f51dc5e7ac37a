"""CPU suite (-m "not gpu"): pins the oracle (cv2-derived fixtures, frozen vectors, numpy second opinion, dense-vs-sparse
self-consistency, analytic invariants), checks the host logic and that the C-ABI library loads and exports its header."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import np_oracle as NP
from oracle import oracle_py as O
from ransac_slam_b200 import synth, sweep
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


# ---- OpenCV primitives restated by the oracle vs cv2's own outputs --------------------------------------------------
def test_remap_matches_cv2_bit_exact():
    fx = np.load(os.path.join(GOLD, "cv_fixtures.npz"))
    for k in range(6):
        got = O.cv_remap(fx[f"remap_src_{k}"], fx[f"remap_mx_{k}"], fx[f"remap_my_{k}"])
        assert np.array_equal(got, fx[f"remap_dst_{k}"].astype(np.float64)), k


def test_covar_matches_cv2():
    fx = np.load(os.path.join(GOLD, "cv_fixtures.npz"))
    for k in range(3):
        M = fx[f"covar_M_{k}"].astype(np.float64)
        cov = fx[f"covar_cov_{k}"]
        nv = M.shape[1]
        ref = cov / np.sqrt(np.outer(np.diag(cov), np.diag(cov)))
        got = O.corrcoef(M, row0_only=False)
        np.testing.assert_allclose(got, ref, rtol=0, atol=5e-15)
        got0 = O.corrcoef(M, row0_only=True)
        np.testing.assert_allclose(got0[0], ref[0], rtol=0, atol=5e-15)
        assert got.shape == (nv, nv)


# ---- frozen oracle vectors (regression) -------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["ref", "noq1"])
def test_oracle_reproduces_frozen_vectors(tag):
    g = np.load(os.path.join(GOLD, "oracle_vectors.npz"))
    N, seed, quirks = int(g[f"{tag}_N"]), int(g[f"{tag}_seed"]), int(g[f"{tag}_quirks"])
    scene, x, P = synth.random_spd_state(N, seed=seed)
    seq = synth.make_sequence(scene, T=1, seed=seed + 5, t0=3)
    o = H.oracle_from(scene, x, P, quirks=quirks, sparse=False, fast_corr=False)
    rc, info = o.frame(seq.images[0], seq.u01[0], predict=False)
    f = o.features()
    assert (f["ic"] == g[f"{tag}_ic"]).all() and (f["li"] == g[f"{tag}_li"]).all() and (f["hi"] == g[f"{tag}_hi"]).all()
    assert [rc, info["hyp_run"], info["best_support"], info["n_hyp"], info["num_ic"]] == list(g[f"{tag}_info"])
    xk, Pk = o.get_state()
    np.testing.assert_allclose(xk, g[f"{tag}_x_hi"], rtol=1e-12, atol=1e-14)
    H.assert_P_close(Pk, g[f"{tag}_P_hi"], rtol=1e-11)


# ---- dense (reference-faithful) vs sparse mode -------------------------------------------------------------------------
@pytest.mark.parametrize("quirks", [O.Q_ALL, O.Q_ALL & ~O.Q1])
def test_dense_vs_sparse_mode(quirks):
    scene, x, P = synth.random_spd_state(30, seed=41)
    seq = synth.make_sequence(scene, T=1, seed=46, t0=3)
    res = []
    for sparse in (False, True):
        o = H.oracle_from(scene, x, P, quirks=quirks, sparse=sparse, fast_corr=sparse)
        o.frame(seq.images[0], seq.u01[0], predict=False)
        res.append((o.features(), o.get_state()))
    (fa, (xa, Pa)), (fb, (xb, Pb)) = res
    for k in ("ic", "li", "hi"):
        assert (fa[k] == fb[k]).all()
    np.testing.assert_allclose(xa, xb, rtol=1e-12, atol=1e-14)
    assert np.abs(Pa - Pb).max() <= 1e-12 * np.abs(Pa).max()


# ---- numpy second opinion ------------------------------------------------------------------------------------------------
def test_h_H_S_against_numpy():
    scene, x, P = synth.random_spd_state(25, seed=51)
    cam = scene.cam.as9()
    o = H.oracle_from(scene, x, P)
    o.search_ic_matches(None)
    f = o.features()
    types = np.zeros(scene.N, int)
    h, vis = NP.predict_h(cam, x, types)
    assert (vis == f["has_h"]).all()
    np.testing.assert_allclose(h[vis], f["h"][vis], rtol=0, atol=1e-10)
    for i in np.flatnonzero(vis)[:10]:
        Hn = NP.jacobian_H(cam, x, types, i, f["h"][i])
        np.testing.assert_allclose(o.H_dense(i), Hn, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(f["S"][i], Hn @ P @ Hn.T + np.eye(2), rtol=1e-10)


def test_jacobian_finite_difference():
    scene, x, P = synth.random_spd_state(12, seed=52)
    cam = scene.cam.as9()
    types = np.zeros(scene.N, int)
    o = H.oracle_from(scene, x, P)
    o.search_ic_matches(None)
    f = o.features()
    i = int(np.flatnonzero(f["has_h"])[0])
    Hd = o.H_dense(i)
    off = 13 + 6 * i
    for c in list(range(3)) + list(range(off, off + 6)):  # position and feature columns (quaternion columns use the reference's own convention)
        e = np.zeros_like(x)
        step = 1e-6 * max(1.0, abs(x[c]))
        e[c] = step
        hp, _ = NP.predict_h(cam, x + e, types)
        hm, _ = NP.predict_h(cam, x - e, types)
        fd = (hp[i] - hm[i]) / (2 * step)
        np.testing.assert_allclose(Hd[:, c], fd, rtol=2e-5, atol=1e-6)


def test_update_against_numpy():
    scene, x, P = synth.random_spd_state(20, seed=53)
    seq = synth.make_sequence(scene, T=1, seed=58, t0=3)
    o = H.oracle_from(scene, x, P, quirks=O.Q_ALL & ~O.Q1)
    o.search_ic_matches(seq.images[0])
    o.ransac_hypotheses(seq.u01[0])
    f = o.features()
    idx = np.flatnonzero(f["li"])
    assert idx.size >= 3
    Hs = np.vstack([o.H_dense(i) for i in idx])
    z = f["z"][idx].reshape(-1)
    h = f["h"][idx].reshape(-1)
    xn, Pn = NP.ekf_update(x, P, Hs, z, h)
    o.update_li()
    xo, Po = o.get_state()
    np.testing.assert_allclose(xo, xn, rtol=1e-10, atol=1e-13)
    H.assert_P_close(Po, Pn, rtol=1e-9)
    w = np.linalg.eigvalsh(0.5 * (Po + Po.T))
    assert w.min() > -1e-12 * w.max()


@pytest.mark.parametrize("q1", [True, False])
def test_support_against_numpy(q1):
    scene, x, P = synth.random_spd_state(30, seed=54)
    seq = synth.make_sequence(scene, T=1, seed=59, t0=3)
    quirks = O.Q_ALL if q1 else O.Q_ALL & ~O.Q1
    o = H.oracle_from(scene, x, P, quirks=quirks)
    o.search_ic_matches(seq.images[0])
    f = o.features()
    ic = np.flatnonzero(f["ic"])
    cam = scene.cam.as9()
    types = np.zeros(scene.N, int)
    # replay the reference loop in numpy over the same uniform draws
    u = seq.u01[0]
    best, n_hyp, run, best_mask = 0, 1000, 0, None
    for i in range(len(u)):
        if not i < n_hyp:
            break
        p = ic[int(np.floor(u[i] * ic.size))]
        sup, inl, _ = NP.hypothesis_support(cam, x, P, types, o.H_dense(p), f["h"][p], f["z"][p], f["z"], f["ic"], q1=q1)
        run = i + 1
        if sup > best:
            best, best_mask = sup, inl
            n_hyp = int(np.ceil(np.log(1 - 0.99) / np.log(1 - sup / ic.size))) if sup < ic.size else 0
            if n_hyp == 0:
                break
        if i > n_hyp:
            break
    rc, info = o.ransac_hypotheses(u)
    assert info["best_support"] == best and info["hyp_run"] == run and info["n_hyp"] == n_hyp
    if best_mask is not None:
        li = o.features()["li"]
        assert (li[f["ic"]] == best_mask).all()


# ---- multi-frame behaviour: the filter tracks, P stays symmetric PSD -----------------------------------------------------
def test_sequence_tracks_and_stays_psd():
    scene = synth.make_scene(N=40, seed=7)
    seq = synth.make_sequence(scene, T=5, seed=8)
    o = H.oracle_from(scene, scene.x0, scene.P0, prior=False, sparse=True, quirks=O.Q_ALL & ~O.Q1)
    for k in range(5):
        rc, info = o.frame(seq.images[k], seq.u01[k])
        assert rc == 0 and info["num_ic"] > 10
        f = o.features()
        assert f["li"].sum() + f["hi"].sum() > 8
        xk, Pk = o.get_state()
        assert np.isfinite(xk).all() and np.isfinite(Pk).all()
        assert abs(np.linalg.norm(xk[3:7]) - 1) < 1e-12
        assert np.abs(Pk - Pk.T).max() <= 1e-14 * np.abs(Pk).max()
        assert np.linalg.eigvalsh(0.5 * (Pk + Pk.T)).min() > -1e-10
    assert (o.features()["times_predicted"] >= 4).all()


def test_edge_cases_no_matches_and_cartesian():
    scene, x, P = synth.random_spd_state(10, seed=61)
    o = H.oracle_from(scene, x, P)
    o.search_ic_matches(np.zeros((scene.cam.nRows, scene.cam.nCols), np.uint8))  # constant image: zero-variance candidates (Q10)
    assert not o.features()["ic"].any()
    rc, info = o.ransac_hypotheses(np.random.default_rng(0).random(100))
    assert rc == 1 and info["num_ic"] == 0  # Q9
    o.update_li()
    xk, Pk = o.get_state()
    assert np.array_equal(xk, x) and np.array_equal(Pk, P)  # no measurements: copy (src/ExtendKF.cpp:635-638)
    o.rescue_hi()
    o.update_hi()
    # uniform draws exhausted before termination -> rc 3
    seq = synth.make_sequence(scene, T=1, seed=66, t0=3)
    o2 = H.oracle_from(scene, x, P)
    o2.search_ic_matches(seq.images[0])
    rc, info = o2.ransac_hypotheses(seq.u01[0][:5])
    assert rc == 3 and info["hyp_run"] == 5


def test_lu_inverse_and_gemm():
    rng = np.random.default_rng(3)
    A = rng.standard_normal((37, 37)) + 6 * np.eye(37)
    np.testing.assert_allclose(O.lu_inverse(A) @ A, np.eye(37), atol=1e-12)
    for (m, k, n, ta, tb) in [(2, 50, 50, False, False), (50, 50, 2, False, True), (130, 77, 41, False, False), (64, 300, 129, True, True), (9, 5, 200, False, True)]:
        a = rng.standard_normal((k, m) if ta else (m, k))
        b = rng.standard_normal((n, k) if tb else (k, n))
        ref = (a.T if ta else a) @ (b.T if tb else b)
        np.testing.assert_allclose(O.dgemm(a, b, ta, tb), ref, rtol=1e-12, atol=1e-12)
    O.set_threads(4)
    a, b = rng.standard_normal((300, 200)), rng.standard_normal((200, 310))
    np.testing.assert_allclose(O.dgemm(a, b), a @ b, rtol=1e-12, atol=1e-12)
    O.set_threads(1)


# ---- product hygiene -----------------------------------------------------------------------------------------------------
def test_c_abi_library_loads_and_exports_header():
    import ctypes

    from ransac_slam_b200 import capi

    hdr = open(os.path.join(ROOT, "include", "rslam.h")).read()
    names = set(re.findall(r"\b(rslam_[A-Za-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 30
    L = capi.load()
    for nme in sorted(names):
        assert hasattr(L, nme), f"{nme} declared in include/rslam.h but not exported"
    assert set(capi.EXPORTS) <= names
    # no GPU here: creation must fail loudly, not fall back
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        with pytest.raises(capi.RslamError):
            capi.Filter(synth.Camera().as9(), 10)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "ransac_slam_b200")
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                assert "oracle" not in txt.lower() or fn == "synth.py" and "oracle" not in txt.lower(), f"{fn} mentions the oracle"


def test_synth_is_deterministic():
    a = synth.make_scene(N=16, seed=5)
    b = synth.make_scene(N=16, seed=5)
    assert np.array_equal(a.P0, b.P0) and np.array_equal(a.templates, b.templates)
    s1 = synth.make_sequence(a, T=2, seed=6)
    s2 = synth.make_sequence(b, T=2, seed=6)
    assert np.array_equal(s1.images, s2.images) and np.array_equal(s1.u01, s2.u01)


# ---- multi-rank host logic on gloo, world_size 2 ---------------------------------------------------------------------------
_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from ransac_slam_b200 import sweep
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
rng = np.random.default_rng(123)           # same stream on every rank: replicated inputs
H = 1001
supports = rng.integers(0, 40, H)
supports[[17, 500, 900]] = 77              # ties: the lowest id must win
b, e = sweep.shard_range(H, world, rank)
key = torch.tensor([sweep.local_key(supports[b:e], b)], dtype=torch.int64)
sweep.allreduce_key(key)
s, i = sweep.decode_key(int(key.item()))
assert (s, i) == (77, 17), (s, i)
fb, fe = sweep.shard_filters(4096, world, rank)
cnt = torch.tensor([fe - fb]); dist.all_reduce(cnt)
assert int(cnt.item()) == 4096
dist.destroy_process_group()
print("ok", rank)
'''


def test_sweep_sharding_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29641")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=120)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_key_packing_properties():
    rng = np.random.default_rng(9)
    sup = rng.integers(0, 1000, 5000)
    k = sweep.local_key(sup, 100)
    s, i = sweep.decode_key(k)
    assert s == sup.max() and i == 100 + int(np.argmax(sup))
    # splitting anywhere and taking the max of the shard keys gives the same winner
    for cut in (1, 777, 4999):
        assert max(sweep.local_key(sup[:cut], 100), sweep.local_key(sup[cut:], 100 + cut)) == k
    assert sweep.local_key([], 0) == 0

#!/usr/bin/env python
"""Randomised GPU-vs-oracle parity sweep: many small random scenes (map sizes 1 .. 70, random quirk masks, random seeds), three frames
each through rslam_frame, match / inlier sets bit for bit and x, P to 1e-9 after every frame (each side re-seeded with the oracle's
state, as tests/test_gpu_parity.py::test_prediction_and_sequence does).  tests/test_gpu_fuzz.py runs a short fixed slice of it.

usage: python tools/fuzz_parity.py [first_case] [n_cases]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_py as O  # noqa: E402
from ransac_slam_b200 import capi, synth  # noqa: E402
from tests import helpers as H  # noqa: E402

QUIRKS = (0x7, 0x6, 0x5, 0x3, 0x0)  # C-ABI masks: Q1 | Q4 | Q6 in every useful combination


def case_params(case):
    rng = np.random.default_rng(90000 + case)
    if case >= 10000:  # larger maps: the multi-CTA rescue (N > 128), the two RANSAC set-up paths (N <= 256 / N > 256), four-panel Cholesky
        N = int(rng.choice([100, 128, 129, 200, 256, 257, 300]))
    else:
        N = int(rng.choice([1, 2, 3, 5, 8, 13, 16, 17, 31, 32, 33, 47, 64, 65, 70]))
    # every third case runs with the patch warp (Tracking::pred_patch_fc) on both sides: smooth 41 x 41 appearances, the 13 x 13 predicted
    # patches warped from them by the device and by the oracle instead of uploaded
    return dict(N=N, quirks=int(rng.choice(QUIRKS)), seed=int(rng.integers(1, 1 << 30)), T=3, warp=(case % 3 == 2))


def run_case(case, verbose=False):
    p = case_params(case)
    scene = synth.make_scene(N=p["N"], seed=p["seed"], texture="smooth" if p["warp"] else "noise")
    seq = synth.make_sequence(scene, T=p["T"], seed=p["seed"] + 1, u01_seed=p["seed"] + 2)
    if p["warp"]:
        o = O.OracleFilter(scene.cam.as9(), std_z=scene.std_z, quirks=p["quirks"] | O.Q11, sparse=bool(case & 1), fast_corr=True, warp_patches=True)
        for i in range(scene.N):
            o.add_feature(0, scene.init_patches[i], None, scene.x0[:3], np.eye(3), scene.uv0[i])
        o.set_state(scene.x0, scene.P0, prior=False)
        g = capi.Filter(scene.cam.as9(), scene.N, quirks=p["quirks"], std_z=scene.std_z)
        g.upload_state(scene.x0, scene.P0, prior=False)
        g.upload_feature_init(scene.init_patches, np.tile(scene.x0[:3], (scene.N, 1)), np.tile(np.eye(3).reshape(1, 9), (scene.N, 1)), scene.uv0)
        g.set_patch_warp(True)
    else:
        o = H.oracle_from(scene, scene.x0, scene.P0, prior=False, sparse=bool(case & 1), quirks=p["quirks"] | O.Q11)
        g = H.gpu_from(scene, scene.x0, scene.P0, prior=False, quirks=p["quirks"])
    stats = dict(ic=0, li=0, hi=0)
    for k in range(p["T"]):
        rc_o, ro = o.frame(seq.images[k], seq.u01[k])
        g.frame(seq.images[k][None], seq.u01[k][None])
        fo, fg = o.features(), g.features()
        for key in ("has_h", "ic", "li", "hi"):
            assert (fo[key] == fg[key]).all(), (case, p, k, key)
        assert (fo["z"][fo["ic"]] == fg["z"][fg["ic"]]).all(), (case, p, k, "z")
        xo, Po = o.get_state()
        xg, Pg = g.download_state()
        H.assert_x_close(xg, xo, what=f"case {case} {p} x frame {k}")
        H.assert_P_close(Pg, Po, what=f"case {case} {p} P frame {k}")
        assert np.array_equal(Pg, Pg.T), (case, p, k, "symmetry")
        rg = g.ransac_result()
        assert rc_o == rg["status"], (case, p, k, "status", rc_o, rg)
        for key in ("hyp_run", "best_support", "n_hyp", "num_ic"):
            assert ro[key] == rg[key], (case, p, k, key, ro, rg)
        g.upload_state(xo, Po)
        for key in stats:
            stats[key] += int(fo[key].sum())
    g.close()
    if verbose:
        print(case, p, stats, flush=True)
    return stats


if __name__ == "__main__":
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    tot = dict(ic=0, li=0, hi=0)
    for c in range(first, first + count):
        s = run_case(c, verbose=True)
        for k in tot:
            tot[k] += s[k]
    print("all", count, "cases agree;", tot)

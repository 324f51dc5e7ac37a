#!/usr/bin/env python
"""Freezes the CPU oracle's outputs on two small seeded single-frame cases into tests/golden/oracle_vectors.npz.
They are regression vectors for the oracle itself (CPU suite) and golden input/output pairs for the CUDA path (GPU suite);
the reference has no golden vectors of its own (parity unpinned)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_py as O  # noqa: E402
from ransac_slam_b200 import synth  # noqa: E402
from tests import helpers as H  # noqa: E402

out = {}
for tag, N, seed, quirks in (("ref", 24, 301, O.Q_ALL), ("noq1", 24, 302, O.Q_ALL & ~O.Q1)):
    scene, x, P = synth.random_spd_state(N, seed=seed)
    seq = synth.make_sequence(scene, T=1, seed=seed + 5, t0=3)
    o = H.oracle_from(scene, x, P, quirks=quirks, sparse=False, fast_corr=False)
    o.search_ic_matches(seq.images[0])
    f1 = o.features()
    rc, info = o.ransac_hypotheses(seq.u01[0])
    f2 = o.features()
    o.update_li()
    x_li, P_li = o.get_state()
    o.rescue_hi()
    f3 = o.features()
    o.update_hi()
    x_hi, P_hi = o.get_state()
    out.update({f"{tag}_N": N, f"{tag}_seed": seed, f"{tag}_quirks": quirks, f"{tag}_h": f1["h"], f"{tag}_S": f1["S"], f"{tag}_ic": f1["ic"],
                f"{tag}_z": f1["z"], f"{tag}_li": f2["li"], f"{tag}_hi": f3["hi"], f"{tag}_info": np.array([rc, info["hyp_run"], info["best_support"], info["n_hyp"], info["num_ic"]]),
                f"{tag}_x_li": x_li, f"{tag}_P_li": P_li, f"{tag}_x_hi": x_hi, f"{tag}_P_hi": P_hi})
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_vectors.npz"), **out)
print("wrote oracle_vectors.npz")

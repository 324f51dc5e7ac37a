#!/usr/bin/env python
"""Small driver for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family of the library on inputs small enough for
the sanitizer's slowdown -- a C2 frame sequence (single-filter latency path, stream launches and CUDA-graph replay), a batch of filters
(batched kernels), a 300-feature filter (k > 256: multi-launch Cholesky, left-looking TRSM, DMMA SYRK), map surgery + FAST, and the
support sweep incl. the library's NCCL path with one rank.  No torch (the sanitizer would instrument its start-up as well).
usage: python tools/sanitize_driver.py [quick]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench as B  # noqa: E402
from ransac_slam_b200 import capi, synth  # noqa: E402


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    # 1. single 100-feature filter, patch warp on: plain stream launches, then graph replay
    scene, seq = B.make_c2(1234, 4)
    for graph in (False, True):
        g = B.new_gpu_filter(scene)
        g.set_graph(graph)
        for k in range(2 if quick else 4):
            g.frame(seq.images[k][None], seq.u01[k][None])
        g.sync()
        print("c2 graph=%d" % graph, B.frame_stats(g), flush=True)
        g.close()
    # 2. batch of filters with different map sizes
    Bn = 3
    scs = [synth.make_scene(N=n_, seed=70 + n_, margin=30, min_sep=18, texture="smooth") for n_ in (100, 64, 37)]
    sqs = [synth.make_sequence(s, T=2, seed=5, n_u01=1000) for s in scs]
    g = capi.Filter(scs[0].cam.as9(), 100, batch=Bn)
    for b in range(Bn):
        g.upload_state(scs[b].x0, scs[b].P0, b=b)
        B.upload_appearance(g, scs[b], b=b)
    g.set_patch_warp(True)
    for k in range(2):
        g.frame(np.stack([sq.images[k] for sq in sqs]), np.stack([sq.u01[k] for sq in sqs]))
    g.sync()
    print("batch", [B.frame_stats(g, b) for b in range(Bn)], flush=True)
    g.close()
    # 3. 300 features: the large-k kernels
    if not quick:
        cam = synth.scaled_camera(2)
        sc = synth.make_scene(N=300, seed=9, cam=cam, margin=30, min_sep=18, motion_scale=0.5, texture="smooth")
        sq = synth.make_sequence(sc, T=2, seed=10, n_u01=4096)
        g = capi.Filter(cam.as9(), 300, quirks=0x6, std_a=0.0035, std_alpha=0.0035)
        g.upload_state(sc.x0, sc.P0)
        B.upload_appearance(g, sc)
        g.set_patch_warp(True)
        for k in range(2):
            g.frame(sq.images[k][None], sq.u01[k][None])
        g.sync()
        print("n300", B.frame_stats(g), flush=True)
        g.close()
    # 4. map surgery + FAST on a small map
    sc, x, P = synth.random_spd_state(12, seed=3)
    g = capi.Filter(sc.cam.as9(), 16)
    g.upload_state(x, P)
    g.upload_patches(sc.templates.astype(np.float64))
    img = synth.background(sc.cam, seed=3)
    g.set_image(img)
    g.map_delete_feature(4)
    g.map_add_feature(np.array([201.0, 77.0]))
    g.map_inversedepth_to_cartesian()
    kp, nkp = g.fast_corner_detect_9(0, 0, sc.cam.nCols, sc.cam.nRows, threshold=20)
    g.map_management(2, 14, np.random.default_rng(1).random(100), reference_indexing=False)
    g.sync()
    print("map", g.N, nkp, flush=True)
    g.close()
    # 5. support sweep (dedupe and brute force) + the NCCL path with one rank
    sc, x, P = synth.random_spd_state(40, seed=91)
    sq = synth.make_sequence(sc, T=1, seed=96, t0=3)
    hyp = np.random.default_rng(4).integers(0, 30, 200).astype(np.int32)
    for dedupe in (True, False):
        g = capi.Filter(sc.cam.as9(), 40, dedupe=dedupe)
        g.upload_state(x, P, prior=True)
        g.upload_patches(sc.templates.astype(np.float64))
        g.set_image(sq.images[0])
        g.search_ic_matches()
        key, mask, pairs = g.support_sweep(hyp)
        if dedupe and os.environ.get("RSLAM_SANITIZE_NCCL", "1") == "1":
            comm = capi.Comm.single_process([0])
            k2, m2, _ = comm.support_sweep([g], hyp)
            assert k2 == key and (m2 == mask).all()
            comm.close()
        print("sweep dedupe=%d" % dedupe, capi.decode_key(key), pairs, flush=True)
        g.close()
    print("sanitize driver done", flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Minimal driver for ncu captures: runs a few frames / sweeps of one workload through the C ABI and exits.
usage: python tools/prof_driver.py {c2|c3|c4|c5} [frames]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench as B  # noqa: E402
import bench_extras as X  # noqa: E402
from ransac_slam_b200 import capi, synth  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    graph = os.environ.get("RSLAM_GRAPH", "0") == "1"
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    if wl == "c2":
        scene, seq = B.make_c2(1234, frames)
        g = B.new_gpu_filter(scene)
        g.set_graph(graph)
        for k in range(frames):
            g.frame(seq.images[k][None], seq.u01[k][None])
        g.sync()
        print("c2 done", g.launches, g.download_pose()[:3])
    elif wl == "c3":
        cam, scene, seq = B.make_c3(frames)
        P0 = synth.assemble_P_torch(scene, dev)
        g = B.new_c3_filter(cam, scene, P0, 0)
        g.set_graph(graph)
        for k in range(frames):
            g.frame(seq.images[k][None], seq.u01[k][None])
        g.sync()
        print("c3 done", g.launches, B.frame_stats(g))
        if os.environ.get("RSLAM_PHASES"):
            import ctypes as C
            out = np.zeros(32)
            g.L.rslam_debug_scratch(g.h, 0, out.ctypes.data_as(C.c_void_p))
            n = max(out[20], 1.0)
            print("k_chol_panel phase cycles per launch (CTA 1): load %.0f update %.0f factor %.0f store %.0f over %d launches" % (out[16] / n, out[17] / n, out[18] / n, out[19] / n, int(n)))
    elif wl == "c4":
        N, H = 5000, 100000
        scene, x, P, z = X.make_c4(dev, N)
        hyp = np.random.Generator(np.random.MT19937(99)).integers(0, N, H).astype(np.int32)
        g = capi.Filter(scene.cam.as9(), N, dedupe=os.environ.get("RSLAM_DEDUPE", "1") == "1")
        xd = torch.from_numpy(x).to(dev)
        g.upload_state_device(xd.data_ptr(), P.data_ptr(), x.size, x.size, N, prior=True)
        g.set_matches(z, np.ones(N, dtype=np.uint8))
        g.search_ic_matches()
        for _ in range(frames):
            key, _, pairs = g.support_sweep(hyp, want_mask=False)
        print("c4 done", capi.decode_key(key), pairs)
    elif wl == "c5":
        Bl = int(os.environ.get("RSLAM_C5_FILTERS", "1024"))
        scene, seq = B.make_c2(1234, frames)
        g = capi.Filter(scene.cam.as9(), 100, batch=Bl)
        g.set_graph(graph)
        n = scene.x0.size
        xd = torch.from_numpy(scene.x0).to(dev)
        Pd = torch.from_numpy(np.ascontiguousarray(scene.P0)).to(dev)
        for b in range(Bl):
            g.upload_state_device(xd.data_ptr(), Pd.data_ptr(), n, n, 100, b=b)
            B.upload_appearance(g, scene, b=b)
        g.set_patch_warp(True)
        for k in range(frames):
            g.frame(seq.images[k][None].repeat(Bl, 0), seq.u01[k][None].repeat(Bl, 0))
        g.sync()
        print("c5 done", g.launches)
    elif wl == "cublas":  # the fp64 GEMM the DMMA tiles are compared with (library probe, not product code)
        nn = 6144
        a = torch.randn(nn, nn, dtype=torch.float64, device=dev)
        b = torch.randn(nn, nn, dtype=torch.float64, device=dev)
        for _ in range(frames):
            c = a @ b
        torch.cuda.synchronize()
        print("cublas done", float(c[0, 0]))
    else:
        raise SystemExit("unknown workload")


if __name__ == "__main__":
    main()

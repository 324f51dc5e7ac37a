"""phase clocks of k_ransac_support (debug build with -DRSLAM_SUP_CLOCKS): python tools/sup_clocks.py path/to/librslam_dbg.so"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ransac_slam_b200 import capi
capi.LIB_PATH = sys.argv[1]
import torch
import bench_extras as X
dev = torch.device("cuda", 0)
N, H = 5000, 100000
scene, x, P, z = X.make_c4(dev, N)
hyp = np.random.Generator(np.random.MT19937(99)).integers(0, N, H).astype(np.int32)
for dedupe in (True, False):
    g = capi.Filter(scene.cam.as9(), N, dedupe=dedupe)
    xd = torch.from_numpy(x).to(dev)
    g.upload_state_device(xd.data_ptr(), P.data_ptr(), x.size, x.size, N, prior=True)
    g.set_matches(z, np.ones(N, dtype=np.uint8))
    g.search_ic_matches()
    out = np.zeros(8, dtype=np.uint64)
    g.support_sweep(hyp, want_mask=False)
    g.L.rslam_debug_sup_clocks(out.ctypes.data_as(C.c_void_p))
    g.support_sweep(hyp, want_mask=False)
    g.L.rslam_debug_sup_clocks(out.ctypes.data_as(C.c_void_p))
    n = float(out[4])
    print("dedupe" if dedupe else "brute", "CTAs", int(n), "cycles per CTA: setup %.0f  phase1 %.0f  barrier %.0f  phase2 %.0f" % tuple(float(v) / n for v in out[:4]))
    g.close()

import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ransac_slam_b200 import capi
capi.LIB_PATH = sys.argv[1]
import bench as B
scene, seq = B.make_c2(1234, 6)
g = B.new_gpu_filter(scene)
g.set_graph(False)
g.L.rslam_debug_scratch.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
for k in range(6):
    g.frame(seq.images[k][None], seq.u01[k][None])
    out = np.zeros(32)
    g.L.rslam_debug_scratch(g.h, 0, out.ctypes.data_as(C.c_void_p))
    print(k, B.frame_stats(g), "cycles load/potrf/trinv/loadpan/solve/trail:", out[16:22].astype(int))

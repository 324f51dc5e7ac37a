import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ransac_slam_b200 import capi, synth
from oracle import oracle_py as O
p = np.load("tests/golden/pgm_frames.npz")
frames = p["frames"]
cam = synth.Camera()
o = O.OracleFilter(cam.as9())
o.initialize_x_and_p()
x, P = o.get_state()
g = capi.Filter(cam.as9(), 60)
g.upload_state(x, P, feat_types=np.zeros(0, np.int32))
g.set_patch_warp(True)
o.set_options(O.Q_ALL, sparse=False, fast_corr=True, warp_patches=True)
rng = np.random.default_rng(11)
um = rng.random(200); ur = rng.random(1000)
g.set_image(frames[0])
print(o.map_management(frames[0], 1, 25, um), g.map_management(1, 25, um))
o.map_reset_flags(); o.ekf_prediction(); o.search_ic_matches(frames[0])
g.begin_frame(); g.ekf_prediction(); g.search_ic_matches()
fo, fg = o.features(), g.features()
print("has_h", (fo["has_h"] == fg["has_h"]).all(), "ic o", fo["ic"].sum(), "g", fg["ic"].sum())
bad = np.nonzero(fo["ic"] != fg["ic"])[0]
print("bad", bad)
pg = g.download_patches()
for i in bad[:4]:
    po = o.patch_matching(i)
    print(i, "h", fo["h"][i], fg["h"][i], "S", fo["S"][i].ravel(), fg["S"][i].ravel(), "z", fo["z"][i], fg["z"][i], "patch maxdiff", np.abs(po - pg[i]).max(), "patch range", po.min(), po.max())

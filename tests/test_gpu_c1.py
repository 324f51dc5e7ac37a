"""Configuration C1 on the device: all 100 frames of the reference's bundled sequence (tests/golden/pgm_frames_all.npz) through the C++ host
classes in System::TrackRunning's order (rslam_replay_pgm), with the libc draws of tests/golden/c1_ref_outputs.npz -- the outputs of the
reference's OWN sources on the same frames and draws (tools/c1_replay.py ref).  The reference's replay is defined for the first 42
frames only: at frame 42 no feature is individually compatible and Tracking::ransac_hypotheses indexes an empty vector;
the library reports status 1 there, skips the updates and carries on."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_bundled_sequence_all_frames():
    import c1_replay as C

    res = C.run_gpu()
    g = np.load(os.path.join(C.GOLD, "c1_ref_outputs.npz"))
    assert res["frames"] == 100
    # every frame on which the reference is defined: feature count, matches, inlier sets (as counts) and the camera state to 1e-9
    assert res["reference_frames"] == 42 and res["frames_in_agreement_with_reference"] == 42, res
    assert res["ic"][42] == 0 and res["li"][42] == 0 and res["hi"][42] == 0  # the frame the reference cannot process
    assert max(res["N"]) < 256 and min(res["N"][1:]) > 0
    print("C1 on the device: %.1f frames/s incl. process start and file IO; %d frames agree with the reference" % (res["value"], res["frames_in_agreement_with_reference"]))

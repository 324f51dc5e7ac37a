"""Generates tests/golden/fast_fixtures.npz: 8-bit test windows and the keypoints cv2.FastFeatureDetector (threshold 100 and 40,
non-maximum suppression on, TYPE_9_16) returns for them, in OpenCV's output order.  Run in the build container (cv2 4.13 from the
wheelhouse); the GPU box only reads the .npz.  The reference calls cv::FAST(im, keypoints, 100, true) on a 61 x 41 window
(src/Map.cpp:233-236, 324-338)."""
import os

import cv2
import numpy as np

rng = np.random.default_rng(2024)
imgs, names = [], []


def blocks(h, w, cell, lo, hi):
    g = rng.integers(lo, hi, (h // cell + 1, w // cell + 1)).astype(np.uint8)
    return np.kron(g, np.ones((cell, cell), np.uint8))[:h, :w]


for k in range(6):  # the reference's window size, blocky high-contrast content so that threshold 100 fires
    imgs.append(blocks(41, 61, int(rng.integers(3, 9)), 0, 256))
    names.append(f"blocks{k}")
for k in range(3):  # smoothed noise: few or no corners at 100
    a = rng.integers(0, 256, (41, 61)).astype(np.float32)
    imgs.append(np.clip(cv2.GaussianBlur(a, (0, 0), 1.0 + k) * 1.5 - 60, 0, 255).astype(np.uint8))
    names.append(f"blur{k}")
imgs.append(rng.integers(0, 256, (41, 61)).astype(np.uint8))
names.append("white_noise")
imgs.append(blocks(240, 320, 7, 0, 256))
names.append("full_frame_blocks")
imgs.append(np.full((41, 61), 128, np.uint8))
names.append("flat")
sq = np.zeros((41, 61), np.uint8)
sq[10:30, 15:45] = 255
imgs.append(sq)
names.append("square")
out = {}
for name, im in zip(names, imgs):
    out[name + "_img"] = im
    for t in (100, 40):
        det = cv2.FastFeatureDetector_create(threshold=t, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        kps = det.detect(im, None)
        out[f"{name}_kp{t}"] = np.array([[int(k.pt[0]), int(k.pt[1])] for k in kps], dtype=np.int32).reshape(-1, 2)
        out[f"{name}_resp{t}"] = np.array([k.response for k in kps], dtype=np.float32)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fast_fixtures.npz"), names=np.array(names), **out)
print({n: (len(out[n + "_kp100"]), len(out[n + "_kp40"])) for n in names}, cv2.__version__)

# --- frames of the reference's bundled sequence (data/images_sequences, 8-bit PGM 320x240): the first 12 files in sorted order, kept
#     as a fixture for the FAST parity test and the ROS-free replay (config C1); with cv2's FAST keypoints at the reference's threshold
import glob

files = sorted(glob.glob("/root/reference/data/images_sequences/*.pgm"))[:12]
frames = np.stack([cv2.imread(f, cv2.IMREAD_UNCHANGED) for f in files])
det = cv2.FastFeatureDetector_create(threshold=100, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
kp = [np.array([[int(k.pt[0]), int(k.pt[1])] for k in det.detect(fr, None)], dtype=np.int32).reshape(-1, 2) for fr in frames]
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "pgm_frames.npz"), frames=frames,
                    names=np.array([os.path.basename(f) for f in files]), **{f"kp{i}": k for i, k in enumerate(kp)})
print("pgm frames", frames.shape, [len(k) for k in kp])

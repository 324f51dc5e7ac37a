#!/usr/bin/env python
"""Generates tests/golden/cv_fixtures.npz with OpenCV's own outputs for the two OpenCV primitives the reference calls on the
hot path (run here, where the Python cv2 4.13 wheel exists; the GPU box never needs cv2):
  * cv2.remap(CV_32F, INTER_LINEAR, BORDER_CONSTANT 0)          -- src/Tracking.cpp:272
  * cv2.calcCovarMatrix(COVAR_NORMAL | COVAR_ROWS) float -> double -- src/Converter.cpp:195
The oracle's restatements (cv_remap_linear_const0, corrcoef_opencv) are pinned against these in tests/test_oracle_cpu.py."""
import os

import cv2
import numpy as np

rng = np.random.default_rng(2024)
out = {}
# remap: 41x41 u8-valued float patch sampled at 13x13 float coordinates incl. out-of-range taps and exact .5/32 ties
for k in range(6):
    src = rng.integers(0, 256, (41, 41)).astype(np.float32)
    cx, cy = rng.uniform(8, 32, 2)
    a = rng.uniform(-0.3, 0.3)
    s = rng.uniform(0.7, 1.4)
    jj, ii = np.meshgrid(np.arange(13) - 6, np.arange(13) - 6)
    mx = (cx + s * (np.cos(a) * jj - np.sin(a) * ii)).astype(np.float32)
    my = (cy + s * (np.sin(a) * jj + np.cos(a) * ii)).astype(np.float32)
    if k == 4:  # walk off the border
        mx += 25
    if k == 5:  # coordinates on exact 1/64 grid -> round-half-even ties in the 1/32 quantisation
        mx = (np.round(mx * 64) / 64).astype(np.float32)
        my = (np.round(my * 64) / 64).astype(np.float32)
    dst = cv2.remap(src, mx, my, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    out[f"remap_src_{k}"] = src
    out[f"remap_mx_{k}"] = mx
    out[f"remap_my_{k}"] = my
    out[f"remap_dst_{k}"] = dst
# calcCovarMatrix: 169 samples (rows) x (ncand+1) variables, float input, double output, unscaled
for k, nv in enumerate((2, 9, 40)):
    M = rng.integers(0, 256, (169, nv)).astype(np.float32)
    M[:, 0] = (rng.integers(0, 256 * 1024, 169) / 1024.0).astype(np.float32)  # remap-like fractional grey levels
    cov, mean = cv2.calcCovarMatrix(M, None, cv2.COVAR_NORMAL | cv2.COVAR_ROWS)
    out[f"covar_M_{k}"] = M
    out[f"covar_cov_{k}"] = cov
    out[f"covar_mean_{k}"] = mean
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "cv_fixtures.npz"), **out)
print("wrote cv_fixtures.npz with", len(out), "arrays; cv2", cv2.__version__)

"""Shared helpers for the parity tests: build an oracle filter and a GPU filter from the same synthetic inputs."""
import numpy as np

from oracle import oracle_py as O
from ransac_slam_b200 import synth


def oracle_from(scene, x, P, quirks=O.Q_ALL, sparse=False, prior=True, fast_corr=True, feat_types=None):
    f = O.OracleFilter(scene.cam.as9(), std_z=scene.std_z, quirks=quirks, sparse=sparse, fast_corr=fast_corr, warp_patches=False)
    N = scene.N
    for i in range(N):
        t = 0 if feat_types is None else int(feat_types[i])
        f.add_feature(t, None, scene.templates[i].astype(np.float64), scene.x0[:3], np.eye(3), scene.uv0[i])
    f.set_state(x, P, prior=prior)
    return f


def gpu_from(scene, x, P, quirks=0x7, prior=True, feat_types=None, max_features=None, batch=1, dedupe=True):
    from ransac_slam_b200 import capi

    g = capi.Filter(scene.cam.as9(), max_features or scene.N, batch=batch, quirks=quirks, std_z=scene.std_z, dedupe=dedupe)
    for b in range(batch):
        g.upload_state(x, P, feat_types=feat_types, b=b, prior=prior)
        g.upload_patches(scene.templates.astype(np.float64), b=b)
    return g


def quirks_o2g(q):
    """oracle quirk mask -> C-ABI quirk mask (Q11 concerns the patch warp, not in the ABI yet)."""
    return q & 0x7


def assert_P_close(Pa, Pb, rtol=1e-9, what="P"):
    scale = np.max(np.abs(np.diag(Pb)))
    atol = 1e-12 * scale
    d = np.abs(Pa - Pb)
    lim = rtol * np.maximum(np.abs(Pa), np.abs(Pb)) + atol
    bad = d > lim
    assert not bad.any(), f"{what}: {bad.sum()} entries differ, max |d| = {d.max():.3e}, max|P| = {np.abs(Pb).max():.3e}, worst rel = {(d / (np.abs(Pb) + atol)).max():.3e}"


def assert_x_close(xa, xb, rtol=1e-9, what="x"):
    d = np.abs(xa - xb)
    lim = rtol * np.maximum(np.abs(xa), np.abs(xb)) + 1e-12
    assert (d <= lim).all(), f"{what}: max |d| = {d.max():.3e} at {int(np.argmax(d - lim))}"


def q1_consistent_state(N, seed=0, cam=None, noise_px=0.3, outlier_frac=0.2, pos_jitter=0.0, n_loose=3):
    """A prior state on which the REFERENCE's support scoring finds real inliers.

    Quirk Q1 (src/Tracking.cpp:448) reads the angles of matched feature j from the stacked POSITION vector: (theta_j, phi_j) :=
    (ri_v[2j], ri_v[2j+1]).  On a generic map that scores ~0 inliers for every hypothesis.  Here the map is built so that those
    entries ARE the features' angles: angles are drawn first, the first 2N entries of the stacked anchors are set to them, the
    landmarks follow as anchor + m(theta, phi) / rho.  All N features are matched (z = projection + noise, a fraction replaced by
    gross outliers), so the reference's RANSAC terminates adaptively with a real low-innovation set.  The last n_loose features have a
    wide angular prior and a match ~1.6 px off: not low-innovation inliers, but inside the rescue gate -> a non-empty hi set.
    Returns (cam, x_k_km1, p_k_km1, z (N x 2), ic (N bool))."""
    cam = cam or synth.Camera()
    rng = np.random.default_rng(seed)
    th = rng.uniform(-0.35, 0.35, N)
    ph = rng.uniform(-0.25, 0.25, N)
    rho = 1.0 / rng.uniform(3.0, 12.0, N)
    stacked = np.zeros(3 * N)
    stacked[0:2 * N:2] = th
    stacked[1:2 * N:2] = ph
    stacked[2 * N:] = rng.uniform(-0.2, 0.2, N)
    anchors = stacked.reshape(N, 3)
    m = np.stack([np.cos(ph) * np.sin(th), -np.sin(ph), np.cos(ph) * np.cos(th)], axis=1)
    landmarks = anchors + m / rho[:, None]
    n = 13 + 6 * N
    x = np.zeros(n)
    r_true = np.array([0.01, -0.02, 0.015])
    q_true = np.array([1.0, 0.004, -0.003, 0.002])
    q_true /= np.linalg.norm(q_true)
    x[0:3] = r_true + rng.normal(0, 0.002, 3)
    q = q_true + np.array([0.0, *rng.normal(0, 0.0005, 3)])
    x[3:7] = q / np.linalg.norm(q)
    x[7:10] = rng.normal(0, 0.004, 3)
    x[10:13] = rng.normal(0, 0.001, 3)
    for i in range(N):
        x[13 + 6 * i:16 + 6 * i] = anchors[i] + pos_jitter * rng.standard_normal(3)
        x[16 + 6 * i] = th[i]
        x[17 + 6 * i] = ph[i]
        x[18 + 6 * i] = rho[i] * (1.0 + 0.03 * rng.standard_normal())
    sd = np.zeros(n)
    sd[0:3] = 0.004
    sd[3:7] = 0.001
    sd[7:10] = 0.01
    sd[10:13] = 0.003
    for i in range(N):
        sd[13 + 6 * i:16 + 6 * i] = 0.002
        sd[16 + 6 * i:18 + 6 * i] = 0.0008
        sd[18 + 6 * i] = 0.06 * x[18 + 6 * i]
    for i in range(N - n_loose, N):
        sd[16 + 6 * i:18 + 6 * i] = 0.012
    U = rng.normal(0, 1.0, (n, 6)) * (0.3 * sd)[:, None]
    P = np.diag(sd**2) + U @ U.T
    P = np.asfortranarray(0.5 * (P + P.T))
    uv, depth = synth.project(cam, r_true, q_true, landmarks)
    assert (depth > 0).all()
    z = np.rint(uv + rng.normal(0, noise_px, uv.shape))
    outl = rng.random(N) < outlier_frac
    z[outl] += rng.choice([-1.0, 1.0], (int(outl.sum()), 2)) * rng.uniform(4, 9, (int(outl.sum()), 2))
    for i in range(N - n_loose, N):
        z[i] = np.rint(uv[i]) + np.array([2.0, -1.0]) * (1 if i % 2 else -1)
    inside = (z[:, 0] > 8) & (z[:, 0] < cam.nCols - 8) & (z[:, 1] > 8) & (z[:, 1] < cam.nRows - 8)
    return cam, x, P, z, inside

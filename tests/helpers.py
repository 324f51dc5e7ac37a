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

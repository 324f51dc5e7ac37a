"""Seeded synthetic inputs for the 1-point-RANSAC EKF path.

Input synthesis only -- none of this is on the measured path.  Camera constants come from the reference's
examples/Monocular/initialize_param.yaml:9-21 as parsed at src/System.cpp:34-58; feature initialisation follows
ExtendKF::hinv (src/ExtendKF.cpp:236-265) and Map::add_a_feature_covariance_inverse_depth (src/Map.cpp:339-400);
the initial camera covariance follows ExtendKF::initialize_x_and_p (src/ExtendKF.cpp:32-55, incl. the skipped P(5,5)).

Deviation from SURVEY.md 8d (documented in DESIGN.md): the truth trajectory is a bounded Lissajous sweep with peak
speed ~0.01 m/frame and ~0.001 rad/frame instead of an unbounded constant-velocity run, so that the whole map stays in
view for all 1000 frames (a straight 10 m run leaves the first camera's frustum and the frames become empty).
"""
from dataclasses import dataclass, field

import numpy as np

EPS = np.finfo(np.float64).eps


@dataclass
class Camera:
    k1: float = 0.06333
    k2: float = 0.01390
    nRows: int = 240
    nCols: int = 320
    Cx: float = 1.7945 / 0.0112
    Cy: float = 1.4433 / 0.0112
    f: float = 2.1735
    dx: float = 0.0112
    dy: float = 0.0112

    def as9(self):
        return np.array([self.k1, self.k2, self.nRows, self.nCols, self.Cx, self.Cy, self.f, self.dx, self.dy], dtype=np.float64)

    @property
    def fku(self):
        return self.f / self.dx

    @property
    def fkv(self):
        return self.f / self.dy


def scaled_camera(scale):
    """Same physical sensor and lens at `scale` x the pixel density (used by the N=2000 config so that 2000 13x13
    patches do not overlap)."""
    c = Camera()
    return Camera(k1=c.k1, k2=c.k2, nRows=c.nRows * scale, nCols=c.nCols * scale, Cx=c.Cx * scale, Cy=c.Cy * scale, f=c.f,
                  dx=c.dx / scale, dy=c.dy / scale)


def q2r(q):
    r, x, y, z = q
    return np.array([[r * r + x * x - y * y - z * z, 2 * (x * y - r * z), 2 * (z * x + r * y)],
                     [2 * (x * y + r * z), r * r - x * x + y * y - z * z, 2 * (y * z - r * x)],
                     [2 * (z * x - r * y), 2 * (y * z + r * x), r * r - x * x - y * y + z * z]])


def distort(cam, uv):
    uv = np.atleast_2d(uv)
    xu = (uv[:, 0] - cam.Cx) * cam.dx
    yu = (uv[:, 1] - cam.Cy) * cam.dy
    ru = np.sqrt(xu * xu + yu * yu)
    rd = ru / (1 + cam.k1 * ru**2 + cam.k2 * ru**4)
    for _ in range(10):
        f = rd + cam.k1 * rd**3 + cam.k2 * rd**5 - ru
        fp = 1 + 3 * cam.k1 * rd**2 + 5 * cam.k2 * rd**4
        rd = rd - f / fp
    D = 1 + cam.k1 * rd**2 + cam.k2 * rd**4
    return np.stack([xu / D / cam.dx + cam.Cx, yu / D / cam.dy + cam.Cy], axis=1)


def undistort(cam, uvd):
    uvd = np.atleast_2d(uvd)
    xd = (uvd[:, 0] - cam.Cx) * cam.dx
    yd = (uvd[:, 1] - cam.Cy) * cam.dy
    rd = np.sqrt(xd * xd + yd * yd)
    D = 1 + cam.k1 * rd**2 + cam.k2 * rd**4
    return np.stack([xd * D / cam.dx + cam.Cx, yd * D / cam.dy + cam.Cy], axis=1)


def jacob_undistort(cam, uvd):
    a = uvd[0] - cam.Cx
    b = uvd[1] - cam.Cy
    rd2 = (a * cam.dx) ** 2 + (b * cam.dy) ** 2
    g = cam.k1 + 2 * cam.k2 * rd2
    c = 1 + cam.k1 * rd2 + cam.k2 * rd2 * rd2
    return np.array([[c + a * g * 2 * a * cam.dx**2, a * g * 2 * b * cam.dy**2], [b * g * 2 * a * cam.dx**2, c + b * g * 2 * b * cam.dy**2]])


def project(cam, r, q, X):
    """Distorted pixel of world points X (m x 3) seen from pose (r, q)."""
    R = q2r(q)
    hc = (X - r) @ R  # rows: R^T (X - r)
    uvu = np.stack([cam.Cx + hc[:, 0] / hc[:, 2] * cam.fku, cam.Cy + hc[:, 1] / hc[:, 2] * cam.fkv], axis=1)
    return distort(cam, uvu), hc[:, 2]


def dRq_times_a_by_dq(q, a):
    q0, q1, q2, q3 = q
    M = [2 * np.array([[q0, -q3, q2], [q3, q0, -q1], [-q2, q1, q0]]), 2 * np.array([[q1, q2, q3], [q2, -q1, -q0], [q3, q0, -q1]]),
         2 * np.array([[-q2, q1, q0], [q1, q2, q3], [-q0, q3, -q2]]), 2 * np.array([[-q3, -q0, q1], [q0, -q3, q2], [q1, q2, q3]])]
    return np.stack([m @ a for m in M], axis=1)


def hinv(cam, uvd, xv, rho0):
    uv = undistort(cam, uvd)[0]
    h = np.array([-(cam.Cx - uv[0]) / cam.fku, -(cam.Cy - uv[1]) / cam.fkv, 1.0])
    n = q2r(xv[3:7]) @ h
    return np.array([xv[0], xv[1], xv[2], np.arctan2(n[0], n[2]), np.arctan2(-n[1], np.hypot(n[0], n[2])), rho0])


def feature_init_jacobians(cam, uvd, xv, reference_fill=False):
    """dy_dxv (6x13) and dy_dhd (6x3) of a new inverse-depth feature (src/Map.cpp:339-386).

    reference_fill: the reference comma-initialises the 3 x 2 dgc_dhu with "1/fku, 0, 0, 0, 1/fkv, 0" (src/Map.cpp:379), which Eigen
    lays out row by row as [1/fku 0; 0 0; 1/fkv 0] (quirk Q16) -- what Map::add_a_feature_covariance_inverse_depth really computes.
    The synthetic scene generator keeps the intended [1/fku 0; 0 1/fkv; 0 0] (any SPD prior is a valid input)."""
    q = xv[3:7]
    R = q2r(q)
    uvu = undistort(cam, uvd)[0]
    Xc = np.array([-(cam.Cx - uvu[0]) / cam.fku, -(cam.Cy - uvu[1]) / cam.fkv, 1.0])
    Xw, Yw, Zw = R @ Xc
    dgw_dq = dRq_times_a_by_dq(q, Xc)
    dth = np.array([Zw / (Xw * Xw + Zw * Zw), 0, -Xw / (Xw * Xw + Zw * Zw)])
    n2 = Xw * Xw + Yw * Yw + Zw * Zw
    s = np.sqrt(Xw * Xw + Zw * Zw)
    dph = np.array([(Xw * Yw) / (n2 * s), -s / n2, (Zw * Yw) / (n2 * s)])
    dy_dq = np.zeros((6, 4))
    dy_dq[3] = dth @ dgw_dq
    dy_dq[4] = dph @ dgw_dq
    dy_dxv = np.zeros((6, 13))
    dy_dxv[0:3, 0:3] = np.eye(3)
    dy_dxv[:, 3:7] = dy_dq
    dyp_dgw = np.zeros((5, 3))
    dyp_dgw[3] = dth
    dyp_dgw[4] = dph
    dgc_dhu = np.array([[1 / cam.fku, 0], [0, 0], [1 / cam.fkv, 0]]) if reference_fill else np.array([[1 / cam.fku, 0], [0, 1 / cam.fkv], [0, 0]])
    dyp_dhd = dyp_dgw @ R @ dgc_dhu @ jacob_undistort(cam, uvd)
    dy_dhd = np.zeros((6, 3))
    dy_dhd[0:5, 0:2] = dyp_dhd
    dy_dhd[5, 2] = 1
    return dy_dxv, dy_dhd


def initial_camera_state(v0=0.0, w0=1e-11, std_v0=0.025, std_w0=0.025):
    x = np.zeros(13)
    x[3] = 1
    x[7:10] = v0
    x[10:13] = w0
    P = np.zeros((13, 13))
    for i in (0, 1, 2, 3, 4, 6):  # index 5 skipped in the reference (Q7)
        P[i, i] = EPS
    P[7:10, 7:10] = np.eye(3) * std_v0**2
    P[10:13, 10:13] = np.eye(3) * std_w0**2
    return x, P


@dataclass
class Scene:
    cam: Camera
    N: int
    landmarks: np.ndarray  # N x 3 world points
    uv0: np.ndarray  # N x 2 integer pixels in the first image
    templates: np.ndarray  # N x 13 x 13 uint8 appearance
    x0: np.ndarray  # 13 + 6N state right after initialisation (x_k_k of frame 0)
    P0: np.ndarray  # matching covariance (None when assemble_P=False; use P_factors + assemble_P_torch)
    std_z: float = 1.0
    seed: int = 1234
    meta: dict = field(default_factory=dict)
    init_patches: np.ndarray = None  # N x 41 x 41 uint8 appearance at initialisation (texture="smooth" only)


def smooth_patches(rng, N, size=41, sigma=1.6):
    """Band-limited, full-contrast texture patches: white noise blurred by a separable Gaussian and stretched to 0..255 per patch.
    Unlike white noise they survive the bilinear resampling of Tracking::pred_patch_fc (src/Tracking.cpp:164-278) with ZNCC > 0.8."""
    r = int(np.ceil(3 * sigma))
    k = np.exp(-0.5 * (np.arange(-r, r + 1) / sigma) ** 2)
    k /= k.sum()
    g = rng.standard_normal((N, size + 2 * r, size + 2 * r))
    g = np.apply_along_axis(lambda v: np.convolve(v, k, mode="valid"), 1, g)
    g = np.apply_along_axis(lambda v: np.convolve(v, k, mode="valid"), 2, g)
    lo = g.min(axis=(1, 2), keepdims=True)
    hi = g.max(axis=(1, 2), keepdims=True)
    return np.rint(255.0 * (g - lo) / (hi - lo)).astype(np.uint8)


def make_scene(N=100, seed=1234, cam=None, depth=(2.0, 20.0), margin=45, rho_rel_err=0.3, std_rho=1.0, std_z=1.0, dense_P=True,
               min_sep=0, assemble_P=True, motion_scale=1.0, texture="noise"):
    """N landmarks in the frustum of the first camera, inverse-depth coded from the first pose."""
    cam = cam or Camera()
    rng = np.random.default_rng(seed)
    uv = np.zeros((N, 2))
    if min_sep > 0:
        # jittered grid so that patches do not overlap
        nx = int(np.floor((cam.nCols - 2 * margin) / min_sep))
        ny = int(np.floor((cam.nRows - 2 * margin) / min_sep))
        assert nx * ny >= N, "image too small for N non-overlapping patches"
        cells = rng.permutation(nx * ny)[:N]
        # pasted appearance is 13 x 13 (noise templates) or 15 x 15 with the matching window one pixel off-centre (smooth texture)
        jit = max(1, min_sep - (16 if texture == "smooth" else 13))
        uv[:, 0] = margin + (cells % nx) * min_sep + rng.integers(0, jit, N)
        uv[:, 1] = margin + (cells // nx) * min_sep + rng.integers(0, jit, N)
    else:
        uv[:, 0] = rng.integers(margin, cam.nCols - margin, N)
        uv[:, 1] = rng.integers(margin, cam.nRows - margin, N)
    d = rng.uniform(depth[0], depth[1], N)
    # motion_scale < 1: a finer-pixel camera watching a proportionally slower motion (keeps the pixel-domain uncertainties of
    # the reference configuration when the pixel density is raised, see scaled_camera)
    xv, Pxv = initial_camera_state(std_v0=0.025 * motion_scale, std_w0=0.025 * motion_scale)
    std_rho = std_rho * motion_scale
    rho_rel_err = rho_rel_err * motion_scale
    n = 13 + 6 * N
    x0 = np.zeros(n)
    x0[:13] = xv
    J = np.zeros((6 * N, 13))
    Padd = np.diag([std_z**2, std_z**2, std_rho**2])
    blocks = np.zeros((N, 6, 6))
    landmarks = np.zeros((N, 3))
    for i in range(N):
        yi = hinv(cam, uv[i], xv, 1.0)
        m = np.array([np.cos(yi[4]) * np.sin(yi[3]), -np.sin(yi[4]), np.cos(yi[4]) * np.cos(yi[3])])
        # depth along the ray measured as distance / |m| = distance
        landmarks[i] = xv[:3] + d[i] * m
        rho_true = 1.0 / d[i]
        yi[5] = rho_true * (1.0 + rho_rel_err * rng.standard_normal())
        if yi[5] <= 0.01 * rho_true:
            yi[5] = 0.01 * rho_true
        x0[13 + 6 * i:19 + 6 * i] = yi
        dy_dxv, dy_dhd = feature_init_jacobians(cam, uv[i], xv)
        J[6 * i:6 * i + 6] = dy_dxv
        blocks[i] = dy_dhd @ Padd @ dy_dhd.T
    templates = rng.integers(0, 256, size=(N, 13, 13), dtype=np.uint8)
    init_patches = None
    if texture == "smooth":
        # what Map::initialize_a_features stores (src/Map.cpp:286-294): the 41 x 41 neighbourhood of the corner; the 13 x 13 matching
        # template is its centre
        init_patches = smooth_patches(np.random.default_rng(seed + 7919), N)
        templates = np.ascontiguousarray(init_patches[:, 14:27, 14:27])
    if not assemble_P:
        sc = Scene(cam=cam, N=N, landmarks=landmarks, uv0=uv, templates=templates, x0=x0, P0=None, std_z=std_z, seed=seed, init_patches=init_patches)
        sc.meta["P_factors"] = (J, Pxv, blocks)
        sc.meta["motion_scale"] = motion_scale
        return sc
    P0 = np.zeros((n, n), order="F")
    P0[:13, :13] = Pxv
    JP = J @ Pxv
    P0[13:, :13] = JP
    P0[:13, 13:] = JP.T
    if dense_P:
        P0[13:, 13:] = JP @ J.T
    for i in range(N):
        s = 13 + 6 * i
        P0[s:s + 6, s:s + 6] = (JP[6 * i:6 * i + 6] @ J[6 * i:6 * i + 6].T) + blocks[i]
    sc = Scene(cam=cam, N=N, landmarks=landmarks, uv0=uv, templates=templates, x0=x0, P0=P0, std_z=std_z, seed=seed, init_patches=init_patches)
    sc.meta["motion_scale"] = motion_scale
    return sc


def truth_pose(t, scale=1.0):
    """Bounded Lissajous truth trajectory; peak speed ~0.0094 m/frame, peak rate ~0.001 rad/frame (times `scale`)."""
    r = scale * np.array([0.3 * np.sin(2 * np.pi * t / 200.0), 0.15 * np.sin(2 * np.pi * t / 140.0), 0.1 * np.sin(2 * np.pi * t / 260.0)])
    ang = scale * 0.03 * np.sin(2 * np.pi * t / 180.0)  # rotation about +y
    q = np.array([np.cos(ang / 2), 0.0, np.sin(ang / 2), 0.0])
    return r, q


def background(cam, seed=7):
    """Low-contrast band-limited texture (box-filtered noise) so that spurious ZNCC > 0.8 is unlikely."""
    rng = np.random.default_rng(seed)
    g = rng.standard_normal((cam.nRows + 8, cam.nCols + 8))
    k = np.ones(9) / 9.0
    g = np.apply_along_axis(lambda v: np.convolve(v, k, mode="same"), 0, g)
    g = np.apply_along_axis(lambda v: np.convolve(v, k, mode="same"), 1, g)
    g = g[4:-4, 4:-4]
    g = (g - g.mean()) / (g.std() + 1e-12)
    return np.clip(110 + 12 * g, 0, 255).astype(np.uint8)


@dataclass
class Sequence:
    scene: Scene
    images: np.ndarray  # T x rows x cols uint8
    z_true: np.ndarray  # T x N x 2 pasted integer pixel (-1 if not pasted)
    outlier: np.ndarray  # T x N bool
    u01: np.ndarray  # T x n_u01 uniforms in [0,1)
    poses: np.ndarray  # T x 7 truth (r, q)


def make_sequence(scene, T=20, noise_px=0.5, outlier_frac=0.2, seed=99, n_u01=1000, u01_seed=42, t0=1):
    """Render T frames: every landmark's template pasted at round(project(truth) + noise); `outlier_frac` of them
    displaced by 1.5..3.5 px in a random direction.  Frame k is the truth pose at time t0 + k."""
    cam = scene.cam
    rng = np.random.default_rng(seed)
    bg = background(cam)
    images = np.zeros((T, cam.nRows, cam.nCols), dtype=np.uint8)
    z_true = -np.ones((T, scene.N, 2), dtype=np.int32)
    outl = np.zeros((T, scene.N), dtype=bool)
    poses = np.zeros((T, 7))
    for k in range(T):
        r, q = truth_pose(t0 + k, scene.meta.get("motion_scale", 1.0))
        poses[k, :3] = r
        poses[k, 3:] = q
        uv, depth = project(cam, r, q, scene.landmarks)
        uv = uv + noise_px * rng.standard_normal(uv.shape)
        is_out = rng.random(scene.N) < outlier_frac
        ang = rng.uniform(0, 2 * np.pi, scene.N)
        mag = rng.uniform(1.5, 3.5, scene.N)
        uv[is_out, 0] += mag[is_out] * np.cos(ang[is_out])
        uv[is_out, 1] += mag[is_out] * np.sin(ang[is_out])
        zi = np.rint(uv).astype(np.int32)
        img = bg.copy()
        # smooth scenes paste a 15 x 15 crop of the 41 x 41 initial appearance: the warped predicted patch carries the reference's
        # one-pixel offset (quirk Q11), so its best match sits one pixel off the pasted centre and needs the surrounding texture
        hp = 7 if scene.init_patches is not None else 6
        for i in range(scene.N):
            x, y = zi[i]
            if depth[i] <= 0 or x - hp < 0 or y - hp < 0 or x + hp + 1 > cam.nCols or y + hp + 1 > cam.nRows:
                continue
            if scene.init_patches is not None:
                img[y - hp:y + hp + 1, x - hp:x + hp + 1] = scene.init_patches[i][20 - hp:21 + hp, 20 - hp:21 + hp]
            else:
                img[y - 6:y + 7, x - 6:x + 7] = scene.templates[i]
            z_true[k, i] = (x, y)
        images[k] = img
        outl[k] = is_out
    u01 = np.random.default_rng(u01_seed).random((T, n_u01))
    return Sequence(scene=scene, images=images, z_true=z_true, outlier=outl, u01=u01, poses=poses)


def random_spd_state(N, seed=0, cam=None, corr_rank=8, corr_scale=0.05, t0=3, rho_err=0.02):
    """A single-frame test state that is CONSISTENT with make_sequence(scene, t0=t0): camera near the truth pose at time t0,
    features near their true inverse depth, and P = structured prior + U U^T (dense cross terms) so that every block of P is
    exercised while the search ellipses stay small.  Returns (scene, x_k_km1, p_k_km1)."""
    cam = cam or Camera()
    margin, sep = 30, 18
    fits = ((cam.nCols - 2 * margin) // sep) * ((cam.nRows - 2 * margin) // sep) >= N
    scene = make_scene(N=N, seed=seed, cam=cam, margin=margin if fits else 45, min_sep=sep if fits else 0)
    rng = np.random.default_rng(seed + 1)
    n = scene.x0.size
    x = scene.x0.copy()
    r, q = truth_pose(t0)
    x[0:3] = r + rng.normal(0, 0.001, 3)
    dq = np.array([0.0, *rng.normal(0, 0.0004, 3)])
    qq = q + dq
    x[3:7] = qq / np.linalg.norm(qq)
    x[7:10] = rng.normal(0, 0.005, 3)
    x[10:13] = rng.normal(0, 0.001, 3)
    d = np.linalg.norm(scene.landmarks - scene.x0[:3], axis=1)
    idx = 13 + 6 * np.arange(N) + 5
    x[idx] = (1.0 / d) * (1.0 + rho_err * rng.standard_normal(N))
    P = scene.P0.copy()
    P[idx, idx] = (2 * rho_err * x[idx]) ** 2
    P[0:3, 0:3] += np.eye(3) * 0.002**2
    P[3:7, 3:7] += np.eye(4) * 0.0008**2
    sc = np.sqrt(np.maximum(np.diag(P), 1e-10)) * corr_scale
    U = rng.normal(0, 1.0, (n, corr_rank)) * sc[:, None]
    P = P + U @ U.T
    P = np.asfortranarray(0.5 * (P + P.T))
    return scene, x, P


def assemble_P_numpy(scene):
    """The same initial covariance as assemble_P_torch(scene, ...) without options, built with element-wise numpy only (no BLAS): the
    camera prior Pxv is diagonal, so [I;J] Pxv [I;J]^T is a fixed-order sum of 13 outer products c c^T -- bit-reproducible on any
    machine, which is what a committed full-size fixture needs.  Returns a Fortran-ordered n x n array."""
    J, Pxv, blocks = scene.meta["P_factors"]
    assert np.count_nonzero(Pxv - np.diag(np.diag(Pxv))) == 0
    N = scene.N
    n = 13 + 6 * N
    Jf = np.zeros((n, 13))
    Jf[:13] = np.eye(13)
    Jf[13:] = J
    P = np.zeros((n, n), order="F")
    tmp = np.empty((n, n), order="F")
    for k in range(13):
        d = Pxv[k, k]
        if d == 0.0:
            continue
        c = np.sqrt(d) * Jf[:, k]
        np.multiply(c[:, None], c[None, :], out=tmp)  # exactly symmetric: c_i c_j == c_j c_i
        P += tmp
    del tmp
    idx = 13 + 6 * np.arange(N)
    bs = 0.5 * (blocks + blocks.transpose(0, 2, 1))
    for a in range(6):
        for b in range(6):
            P[idx + a, idx + b] += bs[:, a, b]
    return P


def assemble_P_torch(scene, device, rho_std_rel=None, x=None, lowrank=0, lowrank_scale=0.05, seed=0):
    """Assemble the initial covariance of a large scene on the GPU with torch (setup plumbing for the N=2000 / 5000-match
    configs: avoids building and copying multi-GB matrices on the host).  P = [I;J] Pxv [I;J]^T + blockdiag(...) (+ U U^T).
    With rho_std_rel the inverse-depth variances are replaced by (rho_std_rel * rho)^2 (a converged map)."""
    import torch

    J, Pxv, blocks = scene.meta["P_factors"]
    N = scene.N
    n = 13 + 6 * N
    Jf = torch.zeros((n, 13), dtype=torch.float64, device=device)
    Jf[:13] = torch.eye(13, dtype=torch.float64, device=device)
    Jf[13:] = torch.from_numpy(J).to(device)
    P = Jf @ torch.from_numpy(Pxv).to(device) @ Jf.T
    blk = torch.from_numpy(blocks).to(device)
    idx = 13 + 6 * torch.arange(N, device=device)
    for a in range(6):
        for b in range(6):
            P[idx + a, idx + b] += blk[:, a, b]
    if rho_std_rel is not None:
        xr = torch.from_numpy(np.asarray(x if x is not None else scene.x0)).to(device)
        P[idx + 5, idx + 5] = (rho_std_rel * xr[idx + 5]) ** 2
    if lowrank > 0:
        g = torch.Generator(device="cpu").manual_seed(seed)
        U = torch.randn((n, lowrank), dtype=torch.float64, generator=g).to(device)
        U = U * (torch.sqrt(torch.clamp(torch.diagonal(P), min=1e-10)) * lowrank_scale)[:, None]
        P += U @ U.T
    P = 0.5 * (P + P.T)
    return P.contiguous()

"""Deterministic synthetic calibration problems (benchmark / test INPUT generator).

Not part of the product path: it only manufactures board poses and noisy corner
observations of the shape `data/calib_example.json` describes, so
that bench.py, the tests and smoke() feed the CUDA engine, the CPU oracle and the
reference build the very same numbers.  Everything is driven by splitmix64
counters, so the data are bit-identical on every machine and numpy version.

Board layout follows unified_calibration.cpp:286-292 (point (size*j, size*i, 0),
rows outer, columns inner).  Ground-truth intrinsics come from
data/ex_epipolar_stereo.json:2,4,6 and the initial guess from
data/calib_example.json:17 of the reference tree.
"""
from __future__ import annotations

import numpy as np

EUCM, UCM, MEI = 0, 1, 2
MODEL_NAMES = {EUCM: "eucm", UCM: "ucm", MEI: "mei"}
NUM_PARAMS = {EUCM: 6, UCM: 5, MEI: 10}

EUCM_GT_LEFT = np.array([0.595728, 0.768828, 307.318, 289.542, 642.617, 398.42])
EUCM_GT_RIGHT = np.array([0.593948, 0.774335, 307.356, 289.482, 637.871, 396.818])
EUCM_GUESS = np.array([0.5, 1.0, 300.0, 300.0, 600.0, 500.0])
STEREO_GT = np.array([0.197255, 0.000222456, -0.00421324, -0.00570702, 0.00103386, -0.0140923])
STEREO_PRIOR = np.array([0.2, 0.0, 0.0, 0.0, 0.0, 0.0])
MEI_GT = np.array([0.9, -0.05, 0.01, 0.002, 0.001, -0.001, 400.0, 400.0, 640.0, 400.0])
MEI_GUESS = np.array([1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 380.0, 380.0, 620.0, 410.0])
UCM_GT = np.array([1.1, 420.0, 415.0, 640.0, 400.0])
UCM_GUESS = np.array([1.0, 400.0, 400.0, 620.0, 410.0])
IMAGE_W, IMAGE_H = 1280, 800

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser over an array of uint64 counters."""
    with np.errstate(over="ignore"):
        z = (x.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def uniform(seed: int, stream: int, n: int, offset: int = 0) -> np.ndarray:
    """n doubles in [0,1), a pure function of (seed, stream, offset + i)."""
    base = splitmix64(np.array([seed * 1000003 + stream], dtype=np.uint64))[0]
    with np.errstate(over="ignore"):
        ctr = (np.arange(offset, offset + n, dtype=np.uint64) * np.uint64(0x2545F4914F6CDD1D) + base) & _M64
    return (splitmix64(ctr) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def normal(seed: int, stream: int, n: int) -> np.ndarray:
    """Box-Muller on the same counters."""
    m = (n + 1) // 2
    u1 = uniform(seed, stream, m)
    u2 = uniform(seed, stream + 1, m)
    rad = np.sqrt(-2.0 * np.log(1.0 - u1))
    out = np.empty(2 * m)
    out[0::2] = rad * np.cos(2.0 * np.pi * u2)
    out[1::2] = rad * np.sin(2.0 * np.pi * u2)
    return out[:n]


def make_board(nx: int = 9, ny: int = 6, size: float = 0.1) -> np.ndarray:
    j, i = np.meshgrid(np.arange(nx), np.arange(ny))
    return np.stack([size * j.ravel(), size * i.ravel(), np.zeros(nx * ny)], axis=1).astype(np.float64)


def rodrigues(rv: np.ndarray) -> np.ndarray:
    """(n,3) rotation vectors -> (n,3,3) rotation matrices."""
    rv = np.atleast_2d(rv)
    th = np.linalg.norm(rv, axis=1)
    safe = np.where(th < 1e-12, 1.0, th)
    u = rv / safe[:, None]
    K = np.zeros((rv.shape[0], 3, 3))
    K[:, 0, 1], K[:, 0, 2] = -u[:, 2], u[:, 1]
    K[:, 1, 0], K[:, 1, 2] = u[:, 2], -u[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -u[:, 1], u[:, 0]
    s, c = np.sin(th)[:, None, None], np.cos(th)[:, None, None]
    return np.eye(3)[None] + s * K + (1.0 - c) * (K @ K)


def project(model: int, p: np.ndarray, X: np.ndarray):
    """Forward projection of (...,3) points; returns (uv, valid)."""
    x, y, z = X[..., 0], X[..., 1], X[..., 2]
    if model == EUCM:
        alpha, beta, fu, fv, u0, v0 = p
        rho = np.sqrt(z * z + beta * (x * x + y * y))
        eta = alpha * rho + (1.0 - alpha) * z
        valid = ~(eta < 1e-3)
        if alpha > 0.5:
            with np.errstate(divide="ignore", invalid="ignore"):
                valid &= ~(z / eta < (alpha - 1.0) / (2.0 * alpha - 1.0))
        with np.errstate(divide="ignore", invalid="ignore"):
            uv = np.stack([fu * x / eta + u0, fv * y / eta + v0], axis=-1)
        return uv, valid
    if model == UCM:
        xi, fu, fv, u0, v0 = p
        rho = np.sqrt(x * x + y * y + z * z)
        d = 1.0 / (z + xi * rho)
        return np.stack([fu * x * d + u0, fv * y * d + v0], axis=-1), (z + xi * rho) > 1e-6
    xi, k1, k2, k3, k4, k5, fu, fv, u0, v0 = p
    rho = np.sqrt(x * x + y * y + z * z)
    d = 1.0 / (z + xi * rho)
    xn, yn = x * d, y * d
    r2 = xn * xn + yn * yn
    D = 1.0 + k1 * r2 + k2 * r2 * r2 + k3 * r2 * r2 * r2
    xd = xn * D + 2.0 * k4 * xn * yn + k5 * (r2 + 2.0 * xn * xn)
    yd = yn * D + 2.0 * k5 * xn * yn + k4 * (r2 + 2.0 * yn * yn)
    return np.stack([fu * xd + u0, fv * yd + v0], axis=-1), (z + xi * rho) > 1e-6


def transform_points(xi: np.ndarray, Xb: np.ndarray) -> np.ndarray:
    """xi (n,6) [t,r], Xb (P,3) -> (n,P,3) points R Xb + t."""
    R = rodrigues(xi[:, 3:])
    return np.einsum("nij,pj->npi", R, Xb) + xi[:, None, :3]


def _sample_poses(seed, n, board, visible_fn, margin=20.0):
    """Rejection-sample n board poses whose corners all satisfy visible_fn."""
    centre = board.mean(axis=0)
    out = np.empty((0, 6))
    rnd = 0
    while out.shape[0] < n:
        m = max(256, 3 * (n - out.shape[0]))
        u = uniform(seed, 100 + rnd, 7 * m).reshape(m, 7)
        rnd += 1
        cz = 0.4 + 0.8 * u[:, 0]
        cx = (2.0 * u[:, 1] - 1.0) * 0.6 * cz
        cy = (2.0 * u[:, 2] - 1.0) * 0.4 * cz
        # rotation vector uniform in a ball of 0.6 rad: direction from a cube sample, radius ~ cbrt
        d = 2.0 * u[:, 3:6] - 1.0
        nd = np.linalg.norm(d, axis=1)
        keep = (nd > 1e-3) & (nd <= 1.0)
        rv = d / np.where(nd < 1e-3, 1.0, nd)[:, None] * (0.6 * np.cbrt(u[:, 6]))[:, None]
        R = rodrigues(rv)
        t = np.stack([cx, cy, cz], axis=1) - np.einsum("nij,j->ni", R, centre)
        xi = np.concatenate([t, rv], axis=1)
        keep &= visible_fn(xi, margin)
        out = np.concatenate([out, xi[keep]], axis=0)
        if rnd > 200:
            raise RuntimeError("pose sampling did not converge")
    return np.ascontiguousarray(out[:n])


def _visible(model, intr, board, margin, extra=None):
    def fn(xi, margin_=margin):
        X = transform_points(xi, board)
        if extra is not None:
            X = extra(X)
        uv, valid = project(model, intr, X)
        ok = valid & (uv[..., 0] > margin_) & (uv[..., 0] < IMAGE_W - margin_) \
            & (uv[..., 1] > margin_) & (uv[..., 1] < IMAGE_H - margin_) & np.isfinite(uv).all(axis=-1)
        return ok.all(axis=1)
    return fn


def make_mono(model: int = EUCM, n_img: int = 20, seed: int = 20241, noise_px: float = 0.1,
              nx: int = 9, ny: int = 6, size: float = 0.1, intr_gt=None, intr_guess=None):
    """Monocular problem (configs C1/C2/C3/C5).  Returns a dict of float64 arrays."""
    board = make_board(nx, ny, size)
    if intr_gt is None:
        intr_gt = {EUCM: EUCM_GT_LEFT, UCM: UCM_GT, MEI: MEI_GT}[model]
    if intr_guess is None:
        intr_guess = {EUCM: EUCM_GUESS, UCM: UCM_GUESS, MEI: MEI_GUESS}[model]
    P = board.shape[0]
    xi_gt = _sample_poses(seed, n_img, board, _visible(model, intr_gt, board, 20.0))
    uv, valid = project(model, intr_gt, transform_points(xi_gt, board))
    assert valid.all()
    obs_clean = np.ascontiguousarray(uv.reshape(n_img, 2 * P))
    obs = obs_clean + noise_px * normal(seed, 7, n_img * 2 * P).reshape(n_img, 2 * P)
    pert = uniform(seed, 9, n_img * 6).reshape(n_img, 6) * 2.0 - 1.0
    xi_init = xi_gt + pert * np.array([0.02, 0.02, 0.02, 0.03, 0.03, 0.03])
    return dict(model=model, K=NUM_PARAMS[model], P=P, n_img=n_img, board=board,
                obs=np.ascontiguousarray(obs), obs_clean=obs_clean,
                intr_gt=np.array(intr_gt, dtype=np.float64), intr_init=np.array(intr_guess, dtype=np.float64),
                xi_gt=xi_gt, xi_init=np.ascontiguousarray(xi_init), width=IMAGE_W, height=IMAGE_H)


def make_stereo(n_pairs: int = 20, seed: int = 20244, noise_px: float = 0.1):
    """Stereo EUCM problem (config C4): camera1 chain [board direct], camera2 chain
    [xiCam12 inverse, board direct] as in data/calib_stereo_example.json:49-54,86-92."""
    board = make_board()
    P = board.shape[0]
    R12 = rodrigues(STEREO_GT[None, 3:])[0]
    t12 = STEREO_GT[:3]

    def to_cam2(X):           # X2 = R12^T (X1 - t12)
        return np.einsum("ji,npj->npi", R12, X - t12)

    vis1 = _visible(EUCM, EUCM_GT_LEFT, board, 20.0)
    vis2 = _visible(EUCM, EUCM_GT_RIGHT, board, 20.0, extra=to_cam2)
    xi_gt = _sample_poses(seed, n_pairs, board, lambda xi, m: vis1(xi, m) & vis2(xi, m))
    X1 = transform_points(xi_gt, board)
    uv1, _ = project(EUCM, EUCM_GT_LEFT, X1)
    uv2, _ = project(EUCM, EUCM_GT_RIGHT, to_cam2(X1))
    nz = normal(seed, 7, 2 * n_pairs * 2 * P).reshape(2, n_pairs, 2 * P)
    pert = uniform(seed, 9, n_pairs * 6).reshape(n_pairs, 6) * 2.0 - 1.0
    xi_init = xi_gt + pert * np.array([0.02, 0.02, 0.02, 0.03, 0.03, 0.03])
    return dict(P=P, n_img=n_pairs, board=board,
                obs1=np.ascontiguousarray(uv1.reshape(n_pairs, 2 * P) + noise_px * nz[0]),
                obs2=np.ascontiguousarray(uv2.reshape(n_pairs, 2 * P) + noise_px * nz[1]),
                obs1_clean=np.ascontiguousarray(uv1.reshape(n_pairs, 2 * P)),
                obs2_clean=np.ascontiguousarray(uv2.reshape(n_pairs, 2 * P)),
                intr1_gt=EUCM_GT_LEFT.copy(), intr2_gt=EUCM_GT_RIGHT.copy(),
                intr1_init=EUCM_GUESS.copy(), intr2_init=EUCM_GUESS.copy(),
                xi12_gt=STEREO_GT.copy(), xi12_init=STEREO_PRIOR.copy(),
                xi_gt=xi_gt, xi_init=np.ascontiguousarray(xi_init), width=IMAGE_W, height=IMAGE_H)


# ---- odometry problem: a robot whose odometry couples consecutive poses -----------------------
def _se3_mat(xi):
    T = np.eye(4)
    T[:3, :3] = rodrigues(np.asarray(xi)[None, 3:])[0]
    T[:3, 3] = np.asarray(xi)[:3]
    return T


def _rotvec(R):
    """log of a rotation matrix (angles well below pi here)."""
    c = np.clip((np.trace(R) - 1.0) / 2.0, -1.0, 1.0)
    th = np.arccos(c)
    w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    return 0.5 * w if th < 1e-9 else w * (th / (2.0 * np.sin(th)))


def _se3_vec(T):
    return np.concatenate([T[:3, 3], _rotvec(T[:3, :3])])


def make_odometry(n: int = 40, seed: int = 20246, noise_px: float = 0.1, odo_noise: float = 0.003,
                  extra_unlinked: int = 0):
    """A camera on a wheeled base looking at one fixed board (the use the reference's "odometry" +
    "transformation_prior" datasets are made for, unified_calibration.cpp:742-829):
        X_cam = xiBaseCam^-1 o xiOdom[i]^-1 o xiWorldBoard (X_board)
    i.e. the chain [xiBaseCam INVERSE (global), xiOdom INVERSE (sequence), xiWorldBoard DIRECT (global)].
    Returns GT / initial values, noisy observations and noisy odometry readings (n x 6)."""
    board = make_board()
    P = board.shape[0]
    R_bc = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])      # camera z along the base's x
    tilt = rodrigues(np.array([[0.05, -0.03, 0.02]]))[0]
    T_bc = np.eye(4); T_bc[:3, :3] = R_bc @ tilt; T_bc[:3, 3] = [0.10, 0.02, 0.30]
    T_wB = np.eye(4); T_wB[:3, :3] = R_bc; T_wB[:3, 3] = [1.0, 0.4, 0.55]
    # unicycle trajectory: speed and yaw rate from the counters, kept within +-0.2 m / +-0.35 rad
    u = uniform(seed, 3, 2 * n).reshape(n, 2)
    T = np.eye(4)
    poses, x, th = [T.copy()], 0.0, 0.0
    for i in range(n - 1):
        v = 0.04 * (2.0 * u[i, 0] - 1.0) - 0.2 * x * 0.2
        w = 0.08 * (2.0 * u[i, 1] - 1.0) - 0.2 * th * 0.2
        T = T @ _se3_mat([v, 0.0, 0.0, 0.0, 0.0, w])
        x, th = T[0, 3], np.arctan2(T[1, 0], T[0, 0])
        poses.append(T.copy())
    xi_odom_gt = np.array([_se3_vec(Tp) for Tp in poses])
    # noisy odometry readings: the true increments perturbed, re-integrated
    nz = normal(seed, 5, 6 * n).reshape(n, 6) * odo_noise
    Tr = np.eye(4)
    readings = [_se3_vec(Tr)]
    for i in range(n - 1):
        inc = np.linalg.inv(poses[i]) @ poses[i + 1]
        Tr = Tr @ inc @ _se3_mat(nz[i] * np.array([1, 1, 0.2, 0.2, 0.2, 1]))
        readings.append(_se3_vec(Tr))
    odom = np.array(readings)
    Xc = np.array([(np.linalg.inv(T_bc) @ np.linalg.inv(Tp) @ T_wB @ np.c_[board, np.ones(P)].T).T[:, :3] for Tp in poses])
    uv, valid = project(EUCM, EUCM_GT_LEFT, Xc)
    assert valid.all() and (uv[..., 0] > 0).all() and (uv[..., 0] < IMAGE_W).all() and (uv[..., 1] > 0).all() and (uv[..., 1] < IMAGE_H).all()
    obs = uv.reshape(n, 2 * P) + noise_px * normal(seed, 7, n * 2 * P).reshape(n, 2 * P)
    xi_bc_gt, xi_wB_gt = _se3_vec(T_bc), _se3_vec(T_wB)
    pert = uniform(seed, 9, 12) * 2.0 - 1.0
    scale = np.array([0.01, 0.01, 0.01, 0.01, 0.01, 0.01])
    return dict(P=P, n_img=n, board=board, obs=np.ascontiguousarray(obs), obs_clean=np.ascontiguousarray(uv.reshape(n, 2 * P)),
                intr_gt=EUCM_GT_LEFT.copy(), intr_init=EUCM_GT_LEFT * np.array([1.01, 0.99, 1.005, 0.995, 1.002, 0.998]),
                xi_odom_gt=xi_odom_gt, odom=np.ascontiguousarray(odom), xi_odom_init=np.ascontiguousarray(odom.copy()),
                xi_bc_gt=xi_bc_gt, xi_bc_init=xi_bc_gt + pert[:6] * scale,
                xi_wB_gt=xi_wB_gt, xi_wB_init=xi_wB_gt + pert[6:] * scale,
                status=[1, 1, 0], err_v=0.05, err_w=0.05, lam=0.01, width=IMAGE_W, height=IMAGE_H)


# ---- synthetic board IMAGES (input of the corner detector) -----------------------------------------
def board_image_intrinsics(width: int, height: int, model: int = EUCM) -> np.ndarray:
    """The intrinsics render_board_image sees the board through: the ground-truth camera scaled to the image size."""
    sc = width / IMAGE_W
    intr = np.array({EUCM: EUCM_GT_LEFT, UCM: UCM_GT, MEI: MEI_GT}[model], dtype=np.float64).copy()
    intr[-4:] *= sc
    intr[-1] = intr[-1] / sc * (height / IMAGE_H)
    intr[-3] = intr[-3] / sc * (height / IMAGE_H)
    return intr


def write_pgm(path: str, img: np.ndarray) -> None:
    with open(path, "wb") as f:
        f.write(b"P5\n# rendered board\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
        f.write(np.ascontiguousarray(img, dtype=np.uint8).tobytes())


def write_png(path: str, img: np.ndarray, filter_type: int = 0) -> None:
    """8-bit grey (H, W) or RGB (H, W, 3), non-interlaced; every line with the given filter (0 none, 1 sub, 2 up)."""
    import struct
    import zlib
    a = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = a.shape[:2]
    ch = 1 if a.ndim == 2 else a.shape[2]
    rows = a.reshape(h, w * ch).astype(np.int16)
    if filter_type == 1:
        f = rows.copy(); f[:, ch:] -= rows[:, :-ch]
    elif filter_type == 2:
        f = rows.copy(); f[1:] -= rows[:-1]
    else:
        f = rows
    raw = b"".join(bytes([filter_type]) + (f[y] & 255).astype(np.uint8).tobytes() for y in range(h))

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)
    with open(path, "wb") as out:
        out.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 0 if ch == 1 else 2, 0, 0, 0)) +
                  chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def render_board_image(width: int = 320, height: int = 240, seed: int = 20250, nx: int = 9, ny: int = 6,
                       model: int = EUCM, supersample: int = 3, noise: float = 2.0):
    """An 8-bit image of the (nx+1) x (ny+1)-square checkerboard whose inner corners are the calibration board,
    seen through a camera model scaled to the image size, with a lighting gradient and sensor noise.  Ray casting per
    sub-pixel (back-projection of the pixel, intersection with the board plane), box-filtered: the edges are
    anti-aliased as a real sensor would see them.  Returns (image uint8 (H, W), true corner positions (P, 2))."""
    board = make_board(nx, ny, 0.1)
    intr = board_image_intrinsics(width, height, model)
    u = uniform(seed, 21, 6)
    rv = (2 * u[:3] - 1) * np.array([0.35, 0.35, 0.5])
    centre = board.mean(axis=0)
    R = rodrigues(rv[None])[0]
    t = np.array([(2 * u[3] - 1) * 0.12, (2 * u[4] - 1) * 0.08, 0.75 + 0.35 * u[5]]) - R @ centre
    uv, _ = project(model, intr, board @ R.T + t)
    ss = supersample
    ys, xs = np.meshgrid((np.arange(height * ss) + 0.5) / ss - 0.5, (np.arange(width * ss) + 0.5) / ss - 0.5, indexing="ij")
    # pixel -> ray (pinhole approximation of the inverse is not enough for a fisheye: invert EUCM in closed form)
    xn, yn = (xs - intr[-2]) / intr[-4], (ys - intr[-1]) / intr[-3]
    if model == EUCM:
        a, b = intr[0], intr[1]
        r2 = xn * xn + yn * yn
        det = np.maximum(1 - (2 * a - 1) * b * r2, 0.0)
        zn = (1 - a * a * b * r2) / ((1 - a) + a * np.sqrt(det))
    else:
        xi = intr[0]
        r2 = xn * xn + yn * yn
        g = np.sqrt(np.maximum(1 + r2 * (1 - xi * xi), 0.0))
        en, ed = -g - xi * r2, xi * xi * r2 - 1
        zn = ed / (ed + xi * en)
    ray = np.stack([xn, yn, zn], axis=-1)
    # intersection with the board plane: points R Xb + t, normal n = R e_z
    n = R[:, 2]
    lam = (n @ t) / np.maximum(ray @ n, 1e-9)
    Xb = (ray * lam[..., None] - t) @ R                    # board coordinates
    sq = 0.1
    ix, iy = np.floor(Xb[..., 0] / sq + 1.0), np.floor(Xb[..., 1] / sq + 1.0)       # squares -1 .. nx, -1 .. ny
    inside = (ix >= 0) & (ix <= nx) & (iy >= 0) & (iy <= ny) & (lam > 0)
    dark = ((ix + iy) % 2 == 0) & inside
    val = np.where(dark, 35.0, np.where(inside, 215.0, 120.0))
    val = val.reshape(height, ss, width, ss).mean(axis=(1, 3))
    light = 0.85 + 0.3 * (np.arange(width)[None, :] / width) * (0.5 + np.arange(height)[:, None] / height)
    nz = normal(seed, 23, width * height).reshape(height, width) * noise
    img = np.clip(np.rint(val * light + nz), 0, 255).astype(np.uint8)
    return img, uv

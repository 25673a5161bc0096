"""visgeom_b200 -- B200-native reprojection-residual engine (calibration hot path of visgeom).

This package is a thin ctypes view of libvisgeom_b200.so (CUDA, sm_100a).  It has
NO CPU implementation: importing works anywhere (so symbols can be inspected), but
every compute call needs the CUDA library and a GPU and raises otherwise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_VARIANT = os.environ.get("VG_VARIANT", "")   # developer builds only (see build.py)
LIB_PATH = os.path.join(_HERE, "libvisgeom_b200%s.so" % ("_" + _VARIANT if _VARIANT else ""))

EUCM, UCM, MEI = 0, 1, 2
TRANSFORM_DIRECT, TRANSFORM_INVERSE = 0, 1
NUM_PARAMS = {EUCM: 6, UCM: 5, MEI: 10}
MODEL_BY_NAME = {"eucm": EUCM, "ucm": UCM, "mei": MEI}   # unified_calibration.cpp:146-178

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


class VisgeomError(RuntimeError):
    pass


class SolveOptions(C.Structure):
    _fields_ = [("max_num_iterations", C.c_int), ("function_tolerance", C.c_double),
                ("gradient_tolerance", C.c_double), ("parameter_tolerance", C.c_double),
                ("initial_radius", C.c_double), ("max_radius", C.c_double), ("min_radius", C.c_double),
                ("min_relative_decrease", C.c_double), ("min_lm_diagonal", C.c_double),
                ("max_lm_diagonal", C.c_double), ("jacobi_scaling", C.c_int),
                ("max_consecutive_invalid", C.c_int), ("verbose", C.c_int), ("reserved", C.c_int)]


class SolveSummary(C.Structure):
    _fields_ = [("iterations", C.c_int), ("num_successful", C.c_int), ("num_unsuccessful", C.c_int),
                ("termination", C.c_int), ("initial_cost", C.c_double), ("final_cost", C.c_double),
                ("seconds_total", C.c_double), ("seconds_evaluate", C.c_double), ("num_evaluations", C.c_int)]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p)

_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library; fail loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VisgeomError(
            f"{LIB_PATH} is missing: build it with `python -m visgeom_b200.build` "
            "(there is no CPU fallback for this engine)")
    L = C.CDLL(LIB_PATH)
    L.vg_last_error.restype = C.c_char_p
    L.vg_launch_count.restype = C.c_ulonglong
    L.vg_model_bounds.argtypes = [C.c_int, C.c_int, c_dp, c_dp]
    L.vg_eval_chain.argtypes = [C.c_int, c_dp, C.c_int, C.c_int, c_dp, c_dp, C.c_int, c_ip, c_ip,
                                C.POINTER(c_dp), c_dp, c_dp, C.POINTER(c_dp), c_dp]
    L.vg_eval_chain_dev.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                    c_ip, c_ip, C.POINTER(C.c_void_p), C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p]
    c_up = C.POINTER(C.c_ubyte)
    L.vg_project_points.argtypes = [C.c_int, c_dp, C.c_longlong, c_dp, c_dp, c_dp, c_dp, c_up]
    L.vg_reconstruct_points.argtypes = [C.c_int, c_dp, C.c_longlong, c_dp, c_dp, c_up]
    L.vg_project_points_dev.argtypes = [C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p]
    L.vg_reconstruct_points_dev.argtypes = [C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    c_fp = C.POINTER(C.c_float)
    L.vg_corner_response.argtypes = [c_up, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, c_fp, c_fp, c_fp, c_fp, c_dp,
                                     C.POINTER(C.c_longlong)]
    L.vg_corner_response_dev.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.vg_host_register.argtypes = [C.c_void_p, C.c_size_t]
    L.vg_host_unregister.argtypes = [C.c_void_p]
    L.vg_detect_pattern.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.vg_detector_host_stages.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                          C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p]
    L.vg_subpixel_evaluate.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p]
    L.vg_subpixel_refine.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]
    L.vg_refine_poses.argtypes = [C.c_int, c_dp, C.c_int, C.c_int, c_dp, c_dp, c_dp, C.c_double, C.POINTER(SolveOptions), c_ip,
                                  c_dp, c_ip]
    if hasattr(L, "vg_problem_create"):
        L.vg_problem_create.restype = C.c_void_p
        L.vg_problem_create.argtypes = [C.c_int]
        L.vg_problem_destroy.argtypes = [C.c_void_p]
        L.vg_problem_add_camera.argtypes = [C.c_void_p, C.c_int, c_dp, C.c_int]
        L.vg_problem_set_bounds.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double]
        L.vg_problem_add_transform.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, c_dp]
        L.vg_problem_add_dataset.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dp, C.c_int, c_dp, c_ip,
                                             C.c_int, c_ip, c_ip]
        L.vg_problem_add_transformation_prior.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dp, c_dp]
        L.vg_problem_add_odometry.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, c_dp]
        L.vg_problem_set_pose_constant.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.vg_problem_set_loss.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.vg_eval_transformation_prior.argtypes = [C.c_int, c_dp, c_dp, c_dp, c_dp, c_dp]
        L.vg_eval_odometry_prior.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, c_dp, c_dp, c_dp, c_dp,
                                             c_dp, c_dp, c_dp]
        L.vg_eval_odometry_cost.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, c_ip, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp,
                                            c_dp, c_dp, c_dp]
        L.vg_problem_peer_export.argtypes = [C.c_void_p, C.c_void_p]
        L.vg_problem_peer_connect.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.vg_problem_peer_inbox.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), c_ip]
        L.vg_problem_peer_connect_local.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), c_ip]
        L.vg_problem_set_peer_timeout.argtypes = [C.c_void_p, C.c_longlong]
        L.vg_visual_cov.argtypes = [C.c_int, c_dp, c_dp, C.c_int, c_dp, C.c_double, C.c_int, c_dp, c_dp]
        L.vg_problem_set_allreduce.argtypes = [C.c_void_p, ALLREDUCE_FN, C.c_void_p, C.c_int, C.c_int]
        L.vg_problem_materialize_jacobians.argtypes = [C.c_void_p, C.c_int]
        L.vg_problem_device_buffer.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p),
                                               C.POINTER(C.c_size_t)]
        L.vg_problem_stream.restype = C.c_void_p
        L.vg_problem_stream.argtypes = [C.c_void_p]
        L.vg_problem_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        L.vg_problem_evaluate_async.argtypes = [C.c_void_p]
        L.vg_problem_fetch_reduced.argtypes = [C.c_void_p, c_dp, c_dp]
        L.vg_problem_solve.argtypes = [C.c_void_p, C.POINTER(SolveOptions), C.POINTER(SolveSummary)]
        L.vg_problem_evaluate.argtypes = [C.c_void_p, c_dp, c_dp]
        L.vg_problem_num_shared.argtypes = [C.c_void_p]
        L.vg_problem_get_camera.argtypes = [C.c_void_p, C.c_int, c_dp]
        L.vg_problem_set_camera.argtypes = [C.c_void_p, C.c_int, c_dp]
        L.vg_problem_get_transform.argtypes = [C.c_void_p, C.c_int, c_dp]
        L.vg_problem_set_transform.argtypes = [C.c_void_p, C.c_int, c_dp]
        L.vg_problem_update_observations.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.vg_problem_update_poses.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.vg_problem_residuals.argtypes = [C.c_void_p, C.c_int, c_dp]
        L.vg_solve_options_default.argtypes = [C.POINTER(SolveOptions)]
    _lib = L
    return L


def _check(rc: int):
    if rc < 0:
        raise VisgeomError(f"visgeom_b200 error {rc}: {lib().vg_last_error().decode()}")
    return rc


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _dp(a):
    return a.ctypes.data_as(c_dp)


def device_count() -> int:
    return lib().vg_device_count()


def launch_count() -> int:
    return int(lib().vg_launch_count())


def hessian_entries(model: int, chain_len: int) -> int:
    return _check(lib().vg_hessian_entries(model, chain_len))


def model_bounds(model: int):
    lo, hi = C.c_double(), C.c_double()
    out = []
    for i in range(NUM_PARAMS[model]):
        _check(lib().vg_model_bounds(model, i, C.byref(lo), C.byref(hi)))
        out.append((lo.value, hi.value))
    return out


def eval_chain(model, intr, board, obs, xi_list, status, is_global,
               want_r=True, want_J=True, want_H=False, out=None):
    """Batched GenericProjectionJac::Evaluate through the host-buffer C ABI.

    obs (n_img, 2P); xi_list[e] is (n_img, 6) for a sequence or (6,) for a global
    transform.  Returns dict(r, J_intr, J_xi[], H) of numpy arrays (Ceres layout).
    out: a dict a previous call returned -- its arrays are written again instead of allocating new ones (Ceres keeps
    its residual / Jacobian buffers across evaluations too).
    """
    K = _check(lib().vg_model_num_params(model))       # "invalid camera model name" comes from the library
    intr = _f64(intr); board = _f64(board); obs = _f64(obs)
    n_img, P, Lc = obs.shape[0], board.shape[0], len(xi_list)
    if Lc > 5:
        _check(lib().vg_eval_chain(model, _dp(intr), n_img, P, _dp(board), _dp(obs), Lc, None, None, None,
                                   None, None, None, None))
    xis = [_f64(x) for x in xi_list]
    st = np.ascontiguousarray(status, dtype=np.int32)
    ig = np.ascontiguousarray(is_global, dtype=np.int32)
    xi_ptrs = (c_dp * Lc)(*[_dp(x) for x in xis])
    if out is not None:
        r, Ja, Je, H = out["r"], out["J_intr"], out["J_xi"], out["H"]
        assert (r is not None) == want_r and (Ja is not None) == want_J and (H is not None) == want_H
    else:
        r = np.empty((n_img, 2 * P)) if want_r else None
        Ja = np.empty((n_img, 2 * P, K)) if want_J else None
        Je = [np.empty((n_img, 2 * P, 6)) for _ in range(Lc)] if want_J else None
        H = np.empty((n_img, hessian_entries(model, Lc))) if want_H else None
    je_ptrs = (c_dp * Lc)(*[_dp(j) for j in Je]) if want_J else None
    _check(lib().vg_eval_chain(model, _dp(intr), n_img, P, _dp(board), _dp(obs), Lc,
                               st.ctypes.data_as(c_ip), ig.ctypes.data_as(c_ip), xi_ptrs,
                               _dp(r) if want_r else None, _dp(Ja) if want_J else None, je_ptrs,
                               _dp(H) if want_H else None))
    return dict(r=r, J_intr=Ja, J_xi=Je, H=H)


def eval_transformation_prior(stiffness, xi_prior, xi, want_J=True):
    """Batched TransformationPrior::Evaluate (calib_cost_functions.cpp:215-228): inputs (n, 6) -> r (n, 6), J (n, 6, 6)."""
    st = _f64(stiffness).reshape(-1, 6); xp = _f64(xi_prior).reshape(-1, 6); x = _f64(xi).reshape(-1, 6)
    n = x.shape[0]
    r = np.empty((n, 6)); J = np.empty((n, 6, 6)) if want_J else None
    _check(lib().vg_eval_transformation_prior(n, _dp(st), _dp(xp), _dp(x), _dp(r), _dp(J) if want_J else None))
    return r, J


def eval_odometry_prior(errV, errW, lam, odom1, odom2, xi1, xi2, want_J=True):
    """Batched OdometryPrior::Evaluate (calib_cost_functions.cpp:177-213): -> r (n, 6), J1, J2 (n, 6, 6)."""
    o1 = _f64(odom1).reshape(-1, 6); o2 = _f64(odom2).reshape(-1, 6)
    a = _f64(xi1).reshape(-1, 6); b = _f64(xi2).reshape(-1, 6)
    n = a.shape[0]
    r = np.empty((n, 6))
    J1 = np.empty((n, 6, 6)) if want_J else None
    J2 = np.empty((n, 6, 6)) if want_J else None
    _check(lib().vg_eval_odometry_prior(n, errV, errW, lam, _dp(o1), _dp(o2), _dp(a), _dp(b), _dp(r),
                                        _dp(J1) if want_J else None, _dp(J2) if want_J else None))
    return r, J1, J2


def eval_odometry_cost(errV, errW, lam, dq_list, intr_prior, xi1, xi2, intr, want_J=True):
    """Batched OdometryCost::Evaluate (odometry_cost_function.cpp:202-267): dq_list[b] = the (m_b, 2) wheel-angle
    increments of block b -> r (n, 6), J1, J2 (n, 6, 6), J3 (n, 6, 3)."""
    a = _f64(xi1).reshape(-1, 6); b = _f64(xi2).reshape(-1, 6)
    n = a.shape[0]
    off = np.zeros(n + 1, dtype=np.int32)
    for k in range(n):
        off[k + 1] = off[k] + len(dq_list[k])
    dq = _f64(np.concatenate([np.asarray(q, dtype=np.float64).reshape(-1, 2) for q in dq_list], axis=0)) if n else np.zeros((0, 2))
    ip = _f64(intr_prior); it = _f64(intr)
    r = np.empty((n, 6))
    J1 = np.empty((n, 6, 6)) if want_J else None
    J2 = np.empty((n, 6, 6)) if want_J else None
    J3 = np.empty((n, 6, 3)) if want_J else None
    _check(lib().vg_eval_odometry_cost(n, errV, errW, lam, off.ctypes.data_as(c_ip), _dp(dq), _dp(ip), _dp(a), _dp(b), _dp(it), _dp(r),
                                       _dp(J1) if want_J else None, _dp(J2) if want_J else None, _dp(J3) if want_J else None))
    return r, J1, J2, J3


def visual_cov(model, intr, xi_board, board, feature_variance, cam_poses):
    """TrajectoryVisualQuality::visualCov (trajectory_generation.cpp:185-206) for n camera poses -> (n, 6, 6)."""
    intr = _f64(intr); xb = _f64(xi_board); board = _f64(board); poses = _f64(cam_poses).reshape(-1, 6)
    out = np.empty((poses.shape[0], 6, 6))
    _check(lib().vg_visual_cov(model, _dp(intr), _dp(xb), board.shape[0], _dp(board), float(feature_variance),
                               poses.shape[0], _dp(poses), _dp(out)))
    return out


def eval_chain_dev(model, intr, board, obs, xi_list, status, is_global, n_img, P,
                   r=None, J_intr=None, J_xi=None, H=None, seq_index=None, stream=0):
    """Device-pointer variant: every array argument is an integer device address
    (e.g. torch.Tensor.data_ptr()); asynchronous on `stream`."""
    Lc = len(xi_list)
    st = np.ascontiguousarray(status, dtype=np.int32)
    ig = np.ascontiguousarray(is_global, dtype=np.int32)
    xi_ptrs = (C.c_void_p * Lc)(*[C.c_void_p(x) for x in xi_list])
    je_ptrs = (C.c_void_p * Lc)(*[C.c_void_p(x) if x else None for x in J_xi]) if J_xi else None
    _check(lib().vg_eval_chain_dev(model, intr, n_img, P, board, obs, Lc,
                                   st.ctypes.data_as(c_ip), ig.ctypes.data_as(c_ip), xi_ptrs,
                                   seq_index, r, J_intr, je_ptrs, H, stream))


def project_points(model, intr, X, want_jacobians=True, uv_init=None):
    """ICamera::projectPointCloud + projectionJacobian + intrinsicJacobian for n points (vg_project_points).
    Returns uv (n,2), ok (n,) bool, dPdX (n,2,3), dPdintr (n,2,K)."""
    L = lib()
    intr, X = _f64(intr), _f64(X).reshape(-1, 3)
    n, K = X.shape[0], NUM_PARAMS[model]
    uv = np.zeros((n, 2)) if uv_init is None else _f64(uv_init).copy()
    ok = np.zeros(n, dtype=np.uint8)
    dx = np.zeros((n, 2, 3)) if want_jacobians else None
    da = np.zeros((n, 2, K)) if want_jacobians else None
    nul = C.cast(None, c_dp)
    _check(L.vg_project_points(model, intr.ctypes.data_as(c_dp), n, X.ctypes.data_as(c_dp), uv.ctypes.data_as(c_dp),
                               dx.ctypes.data_as(c_dp) if want_jacobians else nul,
                               da.ctypes.data_as(c_dp) if want_jacobians else nul,
                               ok.ctypes.data_as(C.POINTER(C.c_ubyte))))
    return uv, ok.astype(bool), dx, da


def reconstruct_points(model, intr, uv, X_init=None):
    """ICamera::reconstructPointCloud for n image points (vg_reconstruct_points): X (n,3), ok (n,) bool."""
    L = lib()
    intr, uv = _f64(intr), _f64(uv).reshape(-1, 2)
    n = uv.shape[0]
    X = np.zeros((n, 3)) if X_init is None else _f64(X_init).copy()
    ok = np.zeros(n, dtype=np.uint8)
    _check(L.vg_reconstruct_points(model, intr.ctypes.data_as(c_dp), n, uv.ctypes.data_as(c_dp), X.ctypes.data_as(c_dp),
                                   ok.ctypes.data_as(C.POINTER(C.c_ubyte))))
    return X, ok.astype(bool)


def refine_poses(model, intr, board, obs, poses, loss_a=25.0, options=None):
    """estimateInitialGrid's refinement as independent per-image solves (vg_refine_poses): returns
    (poses (n, 6), iterations (n,), final_cost (n,), termination (n,))."""
    L = lib()
    intr, board, obs = _f64(intr), _f64(board), _f64(obs)
    x = _f64(poses).copy().reshape(-1, 6)
    n, P = x.shape[0], board.shape[0]
    it = np.zeros(n, dtype=np.int32); term = np.zeros(n, dtype=np.int32); cost = np.zeros(n)
    _check(L.vg_refine_poses(model, _dp(intr), n, P, _dp(board), _dp(obs), _dp(x), float(loss_a),
                             C.byref(options) if options is not None else None, it.ctypes.data_as(c_ip), _dp(cost),
                             term.ctypes.data_as(c_ip)))
    return x, it, cost, term


def corner_response(imgs, sigma1=0.7, sigma2=1.4):
    """CornerDetector::computeResponse for a batch of 8-bit images (n, H, W) or one (H, W) (vg_corner_response).
    Returns dict(resp, gradx, grady, imgrad: float32 arrays of the input's shape; avg (n,), count (n,))."""
    L = lib()
    a = np.ascontiguousarray(imgs, dtype=np.uint8)
    single = a.ndim == 2
    b = a[None] if single else a
    n, h, w = b.shape
    outs = [np.empty((n, h, w), dtype=np.float32) for _ in range(4)]
    avg = np.empty(n); cnt = np.empty(n, dtype=np.int64)
    fp = C.POINTER(C.c_float)
    _check(L.vg_corner_response(b.ctypes.data_as(C.POINTER(C.c_ubyte)), n, w, h, float(sigma1), float(sigma2),
                                *[o.ctypes.data_as(fp) for o in outs], avg.ctypes.data_as(c_dp),
                                cnt.ctypes.data_as(C.POINTER(C.c_longlong))))
    if single:
        outs = [o[0] for o in outs]
    return dict(resp=outs[0], gradx=outs[1], grady=outs[2], imgrad=outs[3], avg=avg, count=cnt)


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def host_register(a):
    """Page-lock a numpy array the caller keeps across calls (vg_host_register): eval_chain(..., out=...) then lets the
    DMA write it directly."""
    _check(lib().vg_host_register(a.ctypes.data_as(C.c_void_p), a.nbytes))


def host_unregister(a):
    _check(lib().vg_host_unregister(a.ctypes.data_as(C.c_void_p)))


def detect_pattern(imgs, nx=9, ny=6, improve=True):
    """CornerDetector(nx, ny, 3, improve).detectPattern for a batch of 8-bit images (n, H, W) or one (H, W)
    (vg_detect_pattern).  Returns (found (n,) bool, corners (n, nx ny, 2)); rows of images without a pattern are NaN."""
    L = lib()
    a = np.ascontiguousarray(imgs, dtype=np.uint8)
    single = a.ndim == 2
    b = a[None] if single else a
    n, h, w = b.shape
    corners = np.full((n, nx * ny, 2), np.nan)
    found = np.zeros(n, dtype=np.uint8)
    _check(L.vg_detect_pattern(_vp(b), n, w, h, nx, ny, int(bool(improve)), _vp(corners), _vp(found)))
    corners[found == 0] = np.nan
    return (bool(found[0]), corners[0]) if single else (found.astype(bool), corners)


def detector_host_stages(img, s1, s2, max_val, max_uv, init_radius, nx=9, ny=6, cap=4096):
    """The detector's host stages of one scale (vg_detector_host_stages; no GPU needed)."""
    L = lib()
    img = np.ascontiguousarray(img, dtype=np.uint8); s1 = np.ascontiguousarray(s1, dtype=np.uint8)
    s2 = np.ascontiguousarray(s2, dtype=np.uint8)
    h, w = img.shape
    mv = np.ascontiguousarray(max_val, dtype=np.float32); muv = np.ascontiguousarray(max_uv, dtype=np.int32).reshape(-1, 2)
    cand = np.zeros((cap, 2), dtype=np.int32); n_cand = C.c_int(0)
    grid = np.zeros((nx * ny, 2), dtype=np.int32); start = np.zeros((nx * ny, 5)); reach = np.zeros(nx * ny)
    rc = L.vg_detector_host_stages(_vp(img), _vp(s1), _vp(s2), w, h, _vp(mv), _vp(muv), len(mv), nx, ny, int(init_radius), _vp(cand),
                                   cap, C.addressof(n_cand), _vp(grid), _vp(start), _vp(reach))
    if rc < 0:
        _check(rc)
    return dict(found=rc == 1, cand=cand[:min(cap, n_cand.value)].copy(), grid=grid, start=start, reach=reach)


def subpixel_evaluate(gradx, grady, prior, length, params):
    """SubpixelCorner::Evaluate for n parameter vectors (vg_subpixel_evaluate) -> (cost (n,), gradient (n, 5))."""
    L = lib()
    gx = np.ascontiguousarray(gradx, dtype=np.float32); gy = np.ascontiguousarray(grady, dtype=np.float32)
    h, w = gx.shape
    x = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, 5); n = x.shape[0]
    pr = np.ascontiguousarray(np.broadcast_to(np.asarray(prior, dtype=np.float64).reshape(-1, 2), (n, 2)))
    ln = np.ascontiguousarray(np.broadcast_to(np.asarray(length, dtype=np.float64).reshape(-1), (n,)))
    cost = np.zeros(n); g = np.zeros((n, 5))
    _check(L.vg_subpixel_evaluate(_vp(gx), _vp(gy), w, h, n, _vp(pr), _vp(ln), _vp(x), _vp(cost), _vp(g)))
    return cost, g


def subpixel_refine(gradx, grady, prior, length, start):
    """improveCorners' minimisation for n corners (vg_subpixel_refine) -> (refined (n, 2), iterations (n,))."""
    L = lib()
    gx = np.ascontiguousarray(gradx, dtype=np.float32); gy = np.ascontiguousarray(grady, dtype=np.float32)
    h, w = gx.shape
    x = np.ascontiguousarray(start, dtype=np.float64).reshape(-1, 5); n = x.shape[0]
    pr = np.ascontiguousarray(np.asarray(prior, dtype=np.float64).reshape(n, 2))
    ln = np.ascontiguousarray(np.asarray(length, dtype=np.float64).reshape(n))
    out = np.zeros((n, 2)); it = np.zeros(n, dtype=np.int32)
    _check(L.vg_subpixel_refine(_vp(gx), _vp(gy), w, h, n, _vp(pr), _vp(ln), _vp(x), _vp(out), _vp(it)))
    return out, it


class Problem:
    """Mirror of what GenericCameraCalibration assembles (unified_calibration.cpp:514-630)
    and hands to ceres::Solve (:39-53), executed on the GPU."""

    def __init__(self, device: int = -1):
        self.L = lib()
        if not hasattr(self.L, "vg_problem_create"):
            raise VisgeomError("library was built without the problem API")
        h = self.L.vg_problem_create(device)
        if not h:
            raise VisgeomError(f"vg_problem_create failed: {self.L.vg_last_error().decode()}")
        self.h = C.c_void_p(h)
        self._K, self._n = {}, {}
        self._cb = None

    def close(self):
        if getattr(self, "h", None):
            self.L.vg_problem_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_camera(self, model, value, constant=False):
        v = _f64(value)
        if len(v) != NUM_PARAMS[model]:
            raise VisgeomError("invalid number of intrinsic parameters")   # unified_calibration.cpp:151,160,169
        cid = _check(self.L.vg_problem_add_camera(self.h, model, _dp(v), int(constant)))
        self._K[cid] = len(v)
        return cid

    def set_bounds(self, cam, idx, lo, hi):
        _check(self.L.vg_problem_set_bounds(self.h, cam, idx, lo, hi))

    def add_transform(self, values, is_global, constant=False):
        v = _f64(values).reshape(-1, 6)
        tid = _check(self.L.vg_problem_add_transform(self.h, int(is_global), int(constant), v.shape[0], _dp(v)))
        self._n[tid] = v.shape[0]
        return tid

    def add_dataset(self, cam, board, obs, transform_ids, status, seq_index=None):
        board = _f64(board); obs = _f64(obs)
        ids = np.ascontiguousarray(transform_ids, dtype=np.int32)
        st = np.ascontiguousarray(status, dtype=np.int32)
        si = None if seq_index is None else np.ascontiguousarray(seq_index, dtype=np.int32)
        return _check(self.L.vg_problem_add_dataset(
            self.h, cam, board.shape[0], _dp(board), obs.shape[0], _dp(obs),
            None if si is None else si.ctypes.data_as(c_ip), len(ids),
            ids.ctypes.data_as(c_ip), st.ctypes.data_as(c_ip)))

    def add_transformation_prior(self, transform, stiffness, index=0, xi_prior=None):
        st = _f64(stiffness)
        xp = None if xi_prior is None else _f64(xi_prior)
        return _check(self.L.vg_problem_add_transformation_prior(self.h, transform, index, _dp(st),
                                                                 None if xp is None else _dp(xp)))

    def add_odometry(self, transform, errV, errW, lam, odom):
        od = _f64(odom).reshape(-1, 6)
        return _check(self.L.vg_problem_add_odometry(self.h, transform, errV, errW, lam, od.shape[0], _dp(od)))

    def set_loss(self, dataset, a):
        """SoftLOneLoss(a) on every block of the dataset (a = 0: no loss)."""
        _check(self.L.vg_problem_set_loss(self.h, dataset, float(a)))

    def set_pose_constant(self, transform, index, constant=True):
        _check(self.L.vg_problem_set_pose_constant(self.h, transform, index, int(constant)))

    def peer_export(self) -> bytes:
        """CUDA IPC handle (64 bytes) of this rank's inbox for the peer-memory exchange."""
        buf = (C.c_ubyte * 64)()
        _check(self.L.vg_problem_peer_export(self.h, buf))
        return bytes(buf)

    def peer_connect(self, rank, nranks, handles: bytes):
        """handles: the nranks x 64 bytes every rank's peer_export() returned, in rank order."""
        assert len(handles) == 64 * nranks
        buf = (C.c_ubyte * len(handles)).from_buffer_copy(handles)
        _check(self.L.vg_problem_peer_connect(self.h, rank, nranks, buf))

    def peer_inbox(self):
        """(device address of this problem's inbox, its device): for problems of one process (peer_connect_local)."""
        ptr, dev = C.c_void_p(), C.c_int()
        _check(self.L.vg_problem_peer_inbox(self.h, C.byref(ptr), C.byref(dev)))
        return ptr.value, dev.value

    def peer_connect_local(self, rank, inboxes):
        """inboxes: every rank's peer_inbox() in rank order (same process)."""
        n = len(inboxes)
        ptrs = (C.c_void_p * n)(*[i[0] for i in inboxes])
        devs = (C.c_int * n)(*[i[1] for i in inboxes])
        _check(self.L.vg_problem_peer_connect_local(self.h, rank, n, ptrs, devs))

    def set_peer_timeout(self, polls):
        _check(self.L.vg_problem_set_peer_timeout(self.h, int(polls)))

    def set_allreduce(self, fn, rank, nranks):
        """fn(buf_ptr:int, count:int, stream:int) -> None sums count doubles in place across ranks."""
        def tramp(ctx, buf, count, stream):
            try:
                fn(buf or 0, count, stream or 0)
                return 0
            except Exception:   # never let an exception cross the C boundary
                import traceback
                traceback.print_exc()
                return -1
        self._cb = ALLREDUCE_FN(tramp)
        _check(self.L.vg_problem_set_allreduce(self.h, self._cb, None, rank, nranks))

    def materialize_jacobians(self, enable=True):
        _check(self.L.vg_problem_materialize_jacobians(self.h, int(enable)))

    def device_buffer(self, dataset, which):
        ptr, nb = C.c_void_p(), C.c_size_t()
        _check(self.L.vg_problem_device_buffer(self.h, dataset, which, C.byref(ptr), C.byref(nb)))
        return ptr.value, nb.value

    @property
    def stream(self):
        return self.L.vg_problem_stream(self.h)

    def set_stream(self, stream_ptr):
        _check(self.L.vg_problem_set_stream(self.h, stream_ptr))

    def evaluate_async(self):
        _check(self.L.vg_problem_evaluate_async(self.h))

    def fetch_reduced(self, want_reduced=True):
        c = C.c_double()
        red = None
        if want_reduced:
            ks = _check(self.L.vg_problem_num_shared(self.h))
            red = np.zeros(ks * ks + ks)
        _check(self.L.vg_problem_fetch_reduced(self.h, C.cast(C.byref(c), c_dp), _dp(red) if red is not None else None))
        return c.value, red

    def default_options(self) -> SolveOptions:
        o = SolveOptions()
        self.L.vg_solve_options_default(C.byref(o))
        return o

    def solve(self, options: SolveOptions | None = None) -> SolveSummary:
        o = options or self.default_options()
        s = SolveSummary()
        _check(self.L.vg_problem_solve(self.h, C.byref(o), C.byref(s)))
        return s

    def evaluate(self, want_reduced=False):
        c = C.c_double()
        red = None
        if want_reduced:
            ks = _check(self.L.vg_problem_num_shared(self.h))
            red = np.zeros(ks * ks + ks)
        _check(self.L.vg_problem_evaluate(self.h, C.cast(C.byref(c), c_dp), _dp(red) if red is not None else None))
        return (c.value, red) if want_reduced else c.value

    def camera(self, cid):
        out = np.zeros(self._K[cid])
        _check(self.L.vg_problem_get_camera(self.h, cid, _dp(out)))
        return out

    def set_camera(self, cid, value):
        v = _f64(value)
        _check(self.L.vg_problem_set_camera(self.h, cid, _dp(v)))

    def transform(self, tid):
        out = np.zeros((self._n[tid], 6))
        _check(self.L.vg_problem_get_transform(self.h, tid, _dp(out)))
        return out

    def set_transform(self, tid, values):
        v = _f64(values).reshape(-1, 6)
        _check(self.L.vg_problem_set_transform(self.h, tid, _dp(v)))

    def set_transform_ptr(self, tid, host_ptr: int):
        """values read from a (pinned) host address"""
        _check(self.L.vg_problem_set_transform(self.h, tid, C.cast(host_ptr, c_dp)))

    def update_poses(self, tid, host_ptr: int):
        """asynchronous upload of a sequence transform's poses from a (pinned) host address"""
        _check(self.L.vg_problem_update_poses(self.h, tid, host_ptr))

    def update_observations(self, dataset, host_ptr: int):
        _check(self.L.vg_problem_update_observations(self.h, dataset, host_ptr))

    def residuals(self, dataset, n_img, P):
        out = np.zeros((n_img, 2 * P))
        _check(self.L.vg_problem_residuals(self.h, dataset, _dp(out)))
        return out

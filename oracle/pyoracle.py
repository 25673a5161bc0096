"""ctypes view of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  Nothing under visgeom_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libvisgeom_oracle.so")
_REF = os.path.join(_HERE, "_ref", "libvisgeom_ref.so")

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("visgeom_oracle.c", "oracle_lm.c", "oracle_tuned.c", "corner_oracle.c", "visgeom_oracle.h",
                                               "oracle_lm.h")]
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libvisgeom_oracle.so"])
    return _LIB


def build_ref(force: bool = False):
    """Compile the reference's own sources against oracle/shim (only where /root/reference exists)."""
    if os.path.isdir("/root/reference/include"):        # make decides whether anything is stale
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []) + ["ref"])
    return _REF if os.path.exists(_REF) else None


class SolveOptions(C.Structure):
    _fields_ = [("max_num_iterations", C.c_int), ("function_tolerance", C.c_double),
                ("gradient_tolerance", C.c_double), ("parameter_tolerance", C.c_double),
                ("initial_radius", C.c_double), ("max_radius", C.c_double), ("min_radius", C.c_double),
                ("min_relative_decrease", C.c_double), ("min_lm_diagonal", C.c_double),
                ("max_lm_diagonal", C.c_double), ("jacobi_scaling", C.c_int),
                ("max_consecutive_invalid", C.c_int), ("verbose", C.c_int), ("threads", C.c_int)]


class SolveSummary(C.Structure):
    _fields_ = [("iterations", C.c_int), ("num_successful", C.c_int), ("num_unsuccessful", C.c_int),
                ("termination", C.c_int), ("initial_cost", C.c_double), ("final_cost", C.c_double),
                ("seconds_total", C.c_double), ("seconds_evaluate", C.c_double), ("num_evaluations", C.c_int)]


def _dp(a):
    return a.ctypes.data_as(c_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class _Evaluator:
    """Shared driver for the two libraries exporting the *_evaluate_batch entry point."""

    def __init__(self, lib, prefix):
        self.lib = lib
        self.prefix = prefix
        fn = getattr(lib, prefix + "_evaluate_batch")
        fn.restype = C.c_int
        fn.argtypes = [C.c_int, c_dp, C.c_int, C.c_int, c_dp, c_dp, C.c_int, c_ip, c_ip,
                       C.POINTER(c_dp), c_dp, c_dp, C.POINTER(c_dp), c_dp, C.c_int]
        self._batch = fn

    def evaluate_batch(self, model, intr, board, obs, xi_list, status, is_global,
                       want_r=True, want_J=True, want_H=False, threads=1):
        """obs (n_img, 2P); xi_list[e] (n_img,6) or (6,).  Returns dict(r, J_intr, J_xi[], H)."""
        K = {0: 6, 1: 5, 2: 10}[model]
        intr = _f64(intr); board = _f64(board); obs = _f64(obs)
        n_img = obs.shape[0]
        P = board.shape[0]
        L = len(xi_list)
        xis = [_f64(x) for x in xi_list]
        st = np.ascontiguousarray(status, dtype=np.int32)
        ig = np.ascontiguousarray(is_global, dtype=np.int32)
        xi_ptrs = (c_dp * L)(*[_dp(x) for x in xis])
        r = np.zeros((n_img, 2 * P)) if want_r else None
        Ja = np.zeros((n_img, 2 * P, K)) if want_J else None
        Je = [np.zeros((n_img, 2 * P, 6)) for _ in range(L)] if want_J else None
        ne = (K + 6 * L + 1) * (K + 6 * L + 2) // 2
        H = np.zeros((n_img, ne)) if want_H else None
        je_ptrs = (c_dp * L)(*[_dp(j) for j in Je]) if want_J else None
        rc = self._batch(model, _dp(intr), n_img, P, _dp(board), _dp(obs), L,
                         st.ctypes.data_as(c_ip), ig.ctypes.data_as(c_ip), xi_ptrs,
                         _dp(r) if want_r else None, _dp(Ja) if want_J else None,
                         je_ptrs, _dp(H) if want_H else None, threads)
        if rc != 0:
            raise RuntimeError(f"{self.prefix}_evaluate_batch failed: {rc}")
        return dict(r=r, J_intr=Ja, J_xi=Je, H=H)


class TransformationPriorC(C.Structure):
    _fields_ = [("xi_prior", C.c_double * 6), ("A", C.c_double * 36), ("R", C.c_double * 9)]


class OdometryPriorC(C.Structure):
    _fields_ = [("zeta_prior", C.c_double * 6), ("A", C.c_double * 36)]


class Oracle(_Evaluator):
    def __init__(self):
        lib = C.CDLL(build())
        super().__init__(lib, "vgo")
        lib.vgo_transformation_prior_init.argtypes = [C.POINTER(TransformationPriorC), c_dp, c_dp]
        lib.vgo_transformation_prior_eval.argtypes = [C.POINTER(TransformationPriorC), c_dp, c_dp, c_dp]
        lib.vgo_odometry_prior_init.argtypes = [C.POINTER(OdometryPriorC), C.c_double, C.c_double, C.c_double, c_dp, c_dp]
        lib.vgo_odometry_prior_eval.argtypes = [C.POINTER(OdometryPriorC), c_dp, c_dp, c_dp, c_dp, c_dp]
        lib.vgo_max_threads.restype = C.c_int
        lib.vgo_lower_bound.restype = C.c_double
        lib.vgo_upper_bound.restype = C.c_double
        lib.vgo_lower_bound.argtypes = [C.c_int, C.c_int]
        lib.vgo_upper_bound.argtypes = [C.c_int, C.c_int]
        for name, n_in in (("vgo_rotation_matrix", 3), ("vgo_inter_omega_rot", 3)):
            getattr(lib, name).argtypes = [c_dp, c_dp]
        for name in ("vgo_compose", "vgo_compose_inverse", "vgo_inverse_compose"):
            getattr(lib, name).argtypes = [c_dp, c_dp, c_dp]
        lib.vgo_project.argtypes = [C.c_int, c_dp, c_dp, c_dp]
        lib.vgo_projection_jacobian.argtypes = [C.c_int, c_dp, c_dp, c_dp, c_dp]
        lib.vgo_intrinsic_jacobian.argtypes = [C.c_int, c_dp, c_dp, c_dp, c_dp]
        lib.vgo_reconstruct.argtypes = [C.c_int, c_dp, c_dp, c_dp]
        lib.vgo_problem_create.restype = C.c_void_p
        lib.vgo_problem_destroy.argtypes = [C.c_void_p]
        lib.vgo_problem_add_camera.argtypes = [C.c_void_p, C.c_int, c_dp, C.c_int]
        lib.vgo_problem_set_bounds.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double]
        lib.vgo_problem_add_transform.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, c_dp]
        lib.vgo_problem_add_dataset.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dp, C.c_int, c_dp, c_ip,
                                                C.c_int, c_ip, c_ip]
        lib.vgo_problem_add_transformation_prior.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dp, c_dp]
        lib.vgo_problem_add_odometry.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, c_dp]
        lib.vgo_problem_set_pose_constant.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        lib.vgo_problem_set_loss.argtypes = [C.c_void_p, C.c_int, C.c_double]
        lib.vgo_problem_solve.argtypes = [C.c_void_p, C.POINTER(SolveOptions), C.POINTER(SolveSummary)]
        lib.vgo_problem_get_camera.argtypes = [C.c_void_p, C.c_int, c_dp]
        lib.vgo_problem_get_transform.argtypes = [C.c_void_p, C.c_int, c_dp]
        lib.vgo_problem_set_camera.argtypes = [C.c_void_p, C.c_int, c_dp]
        lib.vgo_problem_set_transform.argtypes = [C.c_void_p, C.c_int, c_dp]
        lib.vgo_problem_residuals.argtypes = [C.c_void_p, C.c_int, c_dp]
        lib.vgo_problem_evaluate.argtypes = [C.c_void_p, C.c_int, c_dp]
        lib.vgo_solve_options_default.argtypes = [C.POINTER(SolveOptions)]

    def max_threads(self):
        return self.lib.vgo_max_threads()

    def default_options(self) -> SolveOptions:
        o = SolveOptions()
        self.lib.vgo_solve_options_default(C.byref(o))
        return o

    # ---- the prior functors (calib_cost_functions.h:64-108) ----
    def transformation_prior(self, stiffness, xi_prior, xi):
        """-> r (6), J (6, 6): TransformationPrior built from (stiffness, xi_prior), evaluated at xi."""
        st = _f64(stiffness); xp = _f64(xi_prior); x = _f64(xi)
        tp = TransformationPriorC(); r = np.zeros(6); J = np.zeros((6, 6))
        self.lib.vgo_transformation_prior_init(C.byref(tp), _dp(st), _dp(xp))
        self.lib.vgo_transformation_prior_eval(C.byref(tp), _dp(x), _dp(r), _dp(J))
        return r, J

    def odometry_prior(self, errV, errW, lam, odom1, odom2, xi1, xi2):
        """-> r (6), J1, J2 (6, 6): OdometryPrior built from two odometry readings, evaluated at (xi1, xi2)."""
        o1 = _f64(odom1); o2 = _f64(odom2); a = _f64(xi1); b = _f64(xi2)
        op = OdometryPriorC(); r = np.zeros(6); J1 = np.zeros((6, 6)); J2 = np.zeros((6, 6))
        self.lib.vgo_odometry_prior_init(C.byref(op), errV, errW, lam, _dp(o1), _dp(o2))
        self.lib.vgo_odometry_prior_eval(C.byref(op), _dp(a), _dp(b), _dp(r), _dp(J1), _dp(J2))
        return r, J1, J2


    def odometry_cost(self, errV, errW, lam, dq, intr_prior, xi1, xi2, intr):
        """OdometryCost (odometry_cost_function.cpp) -> r, J1, J2 (6 x 6), J3 (6 x 3), zeta_prior, A."""
        q = _f64(dq).reshape(-1, 2); ip = _f64(intr_prior); a = _f64(xi1); b = _f64(xi2); it = _f64(intr)
        oc = OdometryPriorC()
        self.lib.vgo_odometry_cost_init.argtypes = [C.POINTER(OdometryPriorC), C.c_double, C.c_double, C.c_double, C.c_int, c_dp, c_dp]
        self.lib.vgo_odometry_cost_eval.argtypes = [C.POINTER(OdometryPriorC), C.c_int, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]
        if self.lib.vgo_odometry_cost_init(C.byref(oc), errV, errW, lam, q.shape[0], _dp(q), _dp(ip)):
            raise ValueError("OdometryCost needs at least one increment")
        r = np.zeros(6); J1 = np.zeros((6, 6)); J2 = np.zeros((6, 6)); J3 = np.zeros((6, 3))
        self.lib.vgo_odometry_cost_eval(C.byref(oc), q.shape[0], _dp(q), _dp(a), _dp(b), _dp(it), _dp(r), _dp(J1), _dp(J2), _dp(J3))
        return r, J1, J2, J3, np.array(oc.zeta_prior), np.array(oc.A).reshape(6, 6)
    def visual_cov(self, model, intr, xi_board, board, feature_variance, cam_poses):
        """TrajectoryVisualQuality::visualCov for n camera poses -> (n, 6, 6)."""
        intr = _f64(intr); xb = _f64(xi_board); board = _f64(board); poses = _f64(cam_poses).reshape(-1, 6)
        out = np.zeros((poses.shape[0], 6, 6))
        self.lib.vgo_visual_cov.argtypes = [C.c_int, c_dp, c_dp, C.c_int, c_dp, C.c_double, C.c_int, c_dp, c_dp]
        self.lib.vgo_visual_cov(model, _dp(intr), _dp(xb), board.shape[0], _dp(board), feature_variance, poses.shape[0],
                                _dp(poses), _dp(out))
        return out

    # ---- small geometry helpers (for the unit tests of the restatement) ----
    def rotation_matrix(self, v):
        v = _f64(v); R = np.zeros(9)
        self.lib.vgo_rotation_matrix(_dp(v), _dp(R))
        return R.reshape(3, 3)

    def inter_omega_rot(self, v):
        v = _f64(v); R = np.zeros(9)
        self.lib.vgo_inter_omega_rot(_dp(v), _dp(R))
        return R.reshape(3, 3)

    def compose(self, a, b, kind="compose"):
        a = _f64(a); b = _f64(b); o = np.zeros(6)
        getattr(self.lib, "vgo_" + kind)(_dp(a), _dp(b), _dp(o))
        return o

    def project(self, model, params, X):
        params = _f64(params); X = _f64(X); uv = np.zeros(2)
        ok = self.lib.vgo_project(model, _dp(params), _dp(X), _dp(uv))
        return uv, bool(ok)

    def projection_jacobian(self, model, params, X):
        params = _f64(params); X = _f64(X); du = np.zeros(3); dv = np.zeros(3)
        ok = self.lib.vgo_projection_jacobian(model, _dp(params), _dp(X), _dp(du), _dp(dv))
        return np.stack([du, dv]), bool(ok)

    def intrinsic_jacobian(self, model, params, X):
        K = {0: 6, 1: 5, 2: 10}[model]
        params = _f64(params); X = _f64(X); du = np.zeros(K); dv = np.zeros(K)
        ok = self.lib.vgo_intrinsic_jacobian(model, _dp(params), _dp(X), _dp(du), _dp(dv))
        return np.stack([du, dv]), bool(ok)

    def gaussian_blur_u8(self, img, n, sigma):
        """cv::GaussianBlur(img, Size(n, n), sigma, sigma) for an 8-bit image (corner_oracle.c)."""
        img = np.ascontiguousarray(img, dtype=np.uint8)
        out = np.empty_like(img)
        up = C.POINTER(C.c_ubyte)
        self.lib.vgo_gaussian_blur_u8.argtypes = [up, C.c_int, C.c_int, C.c_int, C.c_double, up]
        self.lib.vgo_gaussian_blur_u8(img.ctypes.data_as(up), img.shape[1], img.shape[0], n, float(sigma), out.ctypes.data_as(up))
        return out

    def corner_response(self, img, sigma1, sigma2):
        """CornerDetector::computeResponse (corner_detector.cpp:262-329): dict(resp, gradx, grady, imgrad, avg, count)."""
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w = img.shape
        fp = C.POINTER(C.c_float)
        outs = [np.empty((h, w), dtype=np.float32) for _ in range(4)]
        avg = C.c_double()
        self.lib.vgo_corner_response.restype = C.c_long
        self.lib.vgo_corner_response.argtypes = [C.POINTER(C.c_ubyte), C.c_int, C.c_int, C.c_double, C.c_double, fp, fp, fp, fp,
                                                 C.POINTER(C.c_double)]
        cnt = self.lib.vgo_corner_response(img.ctypes.data_as(C.POINTER(C.c_ubyte)), w, h, float(sigma1), float(sigma2),
                                           *[o.ctypes.data_as(fp) for o in outs], C.byref(avg))
        return dict(resp=outs[0], gradx=outs[1], grady=outs[2], imgrad=outs[3], avg=avg.value, count=int(cnt))

    def reconstruct(self, model, params, uv):
        params = _f64(params); uv = _f64(uv); X = np.zeros(3)
        ok = self.lib.vgo_reconstruct(model, _dp(params), _dp(uv), _dp(X))
        return X, bool(ok)


class OracleProblem:
    """Thin OO wrapper over vgo_problem_* mirroring visgeom_b200.Problem."""

    def __init__(self, oracle: Oracle):
        self.o = oracle
        self.lib = oracle.lib
        self.h = C.c_void_p(self.lib.vgo_problem_create())
        self._K = {}
        self._n = {}

    def close(self):
        if self.h:
            self.lib.vgo_problem_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def add_camera(self, model, value, constant=False):
        v = _f64(value)
        cid = self.lib.vgo_problem_add_camera(self.h, model, _dp(v), int(constant))
        if cid < 0:
            raise ValueError("add_camera failed")
        self._K[cid] = len(v)
        return cid

    def set_bounds(self, cam, idx, lo, hi):
        return self.lib.vgo_problem_set_bounds(self.h, cam, idx, lo, hi)

    def add_transform(self, values, is_global, constant=False):
        v = _f64(values).reshape(-1, 6)
        tid = self.lib.vgo_problem_add_transform(self.h, int(is_global), int(constant), v.shape[0], _dp(v))
        if tid < 0:
            raise ValueError("add_transform failed")
        self._n[tid] = v.shape[0]
        return tid

    def add_dataset(self, cam, board, obs, transform_ids, status, seq_index=None):
        board = _f64(board); obs = _f64(obs)
        ids = np.ascontiguousarray(transform_ids, dtype=np.int32)
        st = np.ascontiguousarray(status, dtype=np.int32)
        si = None if seq_index is None else np.ascontiguousarray(seq_index, dtype=np.int32)
        did = self.lib.vgo_problem_add_dataset(self.h, cam, board.shape[0], _dp(board), obs.shape[0], _dp(obs),
                                               None if si is None else si.ctypes.data_as(c_ip),
                                               len(ids), ids.ctypes.data_as(c_ip), st.ctypes.data_as(c_ip))
        if did < 0:
            raise ValueError(f"add_dataset failed: {did}")
        return did

    def add_transformation_prior(self, transform, stiffness, index=0, xi_prior=None):
        st = _f64(stiffness)
        xp = None if xi_prior is None else _f64(xi_prior)
        rc = self.lib.vgo_problem_add_transformation_prior(self.h, transform, index, _dp(st), None if xp is None else _dp(xp))
        if rc < 0:
            raise ValueError("add_transformation_prior failed")
        return rc

    def add_odometry(self, transform, errV, errW, lam, odom):
        od = _f64(odom).reshape(-1, 6)
        if self.lib.vgo_problem_add_odometry(self.h, transform, errV, errW, lam, od.shape[0], _dp(od)) < 0:
            raise ValueError("add_odometry failed")
        return od.shape[0] - 1

    def set_loss(self, dataset, a):
        if self.lib.vgo_problem_set_loss(self.h, dataset, float(a)) < 0:
            raise ValueError("set_loss failed")

    def set_pose_constant(self, transform, index, constant=True):
        if self.lib.vgo_problem_set_pose_constant(self.h, transform, index, int(constant)) < 0:
            raise ValueError("set_pose_constant failed")

    def solve(self, options=None):
        o = options or self.o.default_options()
        s = SolveSummary()
        rc = self.lib.vgo_problem_solve(self.h, C.byref(o), C.byref(s))
        if rc != 0:
            raise RuntimeError("solve failed")
        return s

    def camera(self, cid):
        out = np.zeros(self._K[cid])
        self.lib.vgo_problem_get_camera(self.h, cid, _dp(out))
        return out

    def transform(self, tid):
        out = np.zeros((self._n[tid], 6))
        self.lib.vgo_problem_get_transform(self.h, tid, _dp(out))
        return out

    def evaluate(self, threads=1):
        c = C.c_double()
        self.lib.vgo_problem_evaluate(self.h, threads, C.cast(C.byref(c), c_dp))
        return c.value


class TunedOracle(_Evaluator):
    """oracle_tuned.c: the same path with shared sub-expressions and no per-image allocation (CPU baseline only)."""

    def __init__(self):
        super().__init__(C.CDLL(build()), "vgo")
        fn = self.lib.vgo_evaluate_batch_tuned
        fn.restype = C.c_int
        fn.argtypes = self._batch.argtypes
        self._batch = fn


class Reference(_Evaluator):
    """The reference's own hot-path sources compiled against oracle/shim (oracle/_ref)."""

    def __init__(self):
        path = build_ref()
        if path is None:
            raise FileNotFoundError("oracle/_ref/libvisgeom_ref.so not built (needs /root/reference)")
        super().__init__(C.CDLL(path), "vgref")
        self.lib.vgref_transformation_prior.argtypes = [c_dp, c_dp, c_dp, c_dp, c_dp]
        self.lib.vgref_odometry_prior.argtypes = [C.c_double, C.c_double, C.c_double, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]

    def visual_cov(self, model, intr, xi_board, nx, ny, step, feature_variance, cam_poses):
        """The reference's own TrajectoryVisualQuality (built through its constructor) -> (n, 6, 6)."""
        intr = _f64(intr); xb = _f64(xi_board); poses = _f64(cam_poses).reshape(-1, 6)
        out = np.zeros((poses.shape[0], 6, 6))
        self.lib.vgref_visual_cov.argtypes = [C.c_int, c_dp, c_dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, c_dp, c_dp]
        self.lib.vgref_visual_cov(model, _dp(intr), _dp(xb), nx, ny, step, feature_variance, poses.shape[0], _dp(poses), _dp(out))
        return out

    def transformation_prior(self, stiffness, xi_prior, xi):
        st = _f64(stiffness); xp = _f64(xi_prior); x = _f64(xi)
        r = np.zeros(6); J = np.zeros((6, 6))
        self.lib.vgref_transformation_prior(_dp(st), _dp(xp), _dp(x), _dp(r), _dp(J))
        return r, J

    def odometry_prior(self, errV, errW, lam, odom1, odom2, xi1, xi2):
        o1 = _f64(odom1); o2 = _f64(odom2); a = _f64(xi1); b = _f64(xi2)
        r = np.zeros(6); J1 = np.zeros((6, 6)); J2 = np.zeros((6, 6))
        self.lib.vgref_odometry_prior(errV, errW, lam, _dp(o1), _dp(o2), _dp(a), _dp(b), _dp(r), _dp(J1), _dp(J2))
        return r, J1, J2


    def odometry_cost(self, errV, errW, lam, dq, intr_prior, xi1, xi2, intr):
        """The reference's own OdometryCost: constructor + Evaluate with three parameter blocks."""
        q = _f64(dq).reshape(-1, 2); ip = _f64(intr_prior); a = _f64(xi1); b = _f64(xi2); it = _f64(intr)
        r = np.zeros(6); J1 = np.zeros((6, 6)); J2 = np.zeros((6, 6)); J3 = np.zeros((6, 3)); zp = np.zeros(6); A = np.zeros((6, 6))
        self.lib.vgref_odometry_cost.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int] + [c_dp] * 11
        self.lib.vgref_odometry_cost(errV, errW, lam, q.shape[0], _dp(q), _dp(ip), _dp(a), _dp(b), _dp(it), _dp(r), _dp(J1), _dp(J2),
                                     _dp(J3), _dp(zp), _dp(A))
        return r, J1, J2, J3, zp, A


class ReferenceDetector:
    """The reference's own checkerboard detector (src/calibration/corner_detector.cpp compiled where it lies against the
    OpenCV / Ceres stand-ins of oracle/shim -> oracle/_ref/libvisgeom_refdet.so; see ref_entry_detector.cpp).
    exact=True loads the -O0 twin, which runs detectPattern as written (improveCorners through initPoin)."""

    def __init__(self, exact: bool = False):
        build_ref()
        path = os.path.join(_HERE, "_ref", "libvisgeom_refdet_O0.so" if exact else "libvisgeom_refdet.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path + " not built (needs /root/reference)")
        self.exact = exact
        self.lib = C.CDLL(path)
        vp = C.c_void_p
        self.lib.vgref_detect_pattern.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp]
        self.lib.vgref_detector_stages.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, vp, vp, vp, vp,
                                                   C.c_int, vp, vp, vp]
        self.lib.vgref_subpixel_evaluate.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_int, C.c_double, vp, vp, vp]
        self.lib.vgref_subpixel_solve.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_int, C.c_double, vp, vp]
        self.lib.vgref_detector_maps.argtypes = [vp, C.c_int, C.c_int, C.c_double, vp, vp, vp, vp, vp, vp]

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(C.c_void_p)

    def detect_pattern(self, img, nx=9, ny=6, improve=True):
        """detectPattern -> (found, corners (nx ny, 2), start values (nx ny, 5) or None, iterations or None)."""
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w = img.shape
        c = np.zeros((nx * ny, 2)); init = np.zeros((nx * ny, 5)); it = np.zeros(nx * ny, dtype=np.int32)
        own = 0 if self.exact else 1
        ok = self.lib.vgref_detect_pattern(self._p(img), w, h, nx, ny, int(bool(improve)), own, self._p(c), self._p(init), self._p(it))
        return bool(ok), c, (init if own and improve else None), (it if own and improve else None)

    def stages(self, img, sigma2, nx=9, ny=6, cap=4096):
        """One scale of detectPattern: candidates in graph order, arcs (per candidate: neighbours, signs), pattern."""
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w = img.shape
        pts = np.zeros((cap, 2), dtype=np.int32); n_arcs = np.zeros(cap, dtype=np.int32)
        arc_cap = 16 * cap
        arcs = np.zeros(arc_cap, dtype=np.int32); sign = np.zeros(arc_cap, dtype=np.int32)
        pattern = np.zeros(nx * ny + 8, dtype=np.int32); n_pat = C.c_int(0); avg = C.c_double(0)
        n = self.lib.vgref_detector_stages(self._p(img), w, h, nx, ny, float(sigma2), cap, self._p(pts), self._p(n_arcs), self._p(arcs),
                                           self._p(sign), arc_cap, self._p(pattern), C.addressof(n_pat), C.addressof(avg))
        n = min(n, cap)
        tot = int(n_arcs[:n].sum())
        return dict(cand=pts[:n].copy(), n_arcs=n_arcs[:n].copy(), arcs=arcs[:tot].copy(), sign=sign[:tot].copy(),
                    pattern=pattern[:n_pat.value].copy(), avg=avg.value)

    def maps(self, img, sigma2):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w = img.shape
        f = [np.zeros((h, w), dtype=np.float32) for _ in range(4)]
        s = [np.zeros((h, w), dtype=np.uint8) for _ in range(2)]
        self.lib.vgref_detector_maps(self._p(img), w, h, float(sigma2), *[self._p(a) for a in f + s])
        return dict(resp=f[0], gradx=f[1], grady=f[2], imgrad=f[3], s1=s[0], s2=s[1])

    def subpixel_evaluate(self, gradx, grady, prior, length, params, steps=7):
        gx = np.ascontiguousarray(gradx, dtype=np.float32); gy = np.ascontiguousarray(grady, dtype=np.float32)
        h, w = gx.shape
        pr = _f64(prior); x = _f64(params)
        cost = C.c_double(0); g = np.zeros(5)
        self.lib.vgref_subpixel_evaluate(self._p(gx), self._p(gy), w, h, self._p(pr), steps, float(length), self._p(x),
                                         C.addressof(cost), self._p(g))
        return cost.value, g

    def subpixel_solve(self, gradx, grady, prior, length, start, steps=7):
        gx = np.ascontiguousarray(gradx, dtype=np.float32); gy = np.ascontiguousarray(grady, dtype=np.float32)
        h, w = gx.shape
        pr = _f64(prior); x = _f64(start).copy()
        cost = C.c_double(0)
        it = self.lib.vgref_subpixel_solve(self._p(gx), self._p(gy), w, h, self._p(pr), steps, float(length), self._p(x), C.addressof(cost))
        return x, it, cost.value

"""Generate the committed golden vectors from the REFERENCE's own code.

Run in the authoring container only (needs /root/reference):
    python tests/golden/make_golden.py
It builds oracle/_ref (the reference's headers + calib_cost_functions.cpp compiled where they lie
against the stand-in Eigen/Ceres headers of oracle/shim -- Eigen and Ceres themselves are not
installed) and records inputs and GenericProjectionJac::Evaluate outputs
(calib_cost_functions.cpp:28-117) for a set of small cases, plus single-function probes of
rotationMatrix / interOmegaRot / Transformation::compose* / reconstructPoint / bounds.
The reference cannot travel to the GPU box, these fixtures can."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import synthdata as sd  # noqa: E402
from oracle.pyoracle import Oracle, Reference, c_dp  # noqa: E402

D, I = 0, 1


def dp(a):
    return a.ctypes.data_as(c_dp)


def cases(oracle):
    out = {}
    for model, name in ((sd.EUCM, "eucm"), (sd.UCM, "ucm"), (sd.MEI, "mei")):
        d = sd.make_mono(model, 4, seed=900 + model)
        out[f"mono_{name}_gt"] = (model, d["intr_gt"], d["board"], d["obs"], [d["xi_gt"]], [D], [0])
        out[f"mono_{name}_init"] = (model, d["intr_init"], d["board"], d["obs"], [d["xi_init"]], [D], [0])
    s = sd.make_stereo(4, seed=910)
    out["stereo_cam2"] = (sd.EUCM, s["intr2_init"], s["board"], s["obs2"], [s["xi12_init"], s["xi_init"]], [I, D], [1, 0])
    # longer mixed chains (built like tests/test_eval_gpu.py::make_chain)
    sys.path.insert(0, os.path.dirname(HERE))
    from test_eval_gpu import make_chain
    for model, name in ((sd.EUCM, "eucm"), (sd.MEI, "mei")):
        d = sd.make_mono(model, 3, seed=920 + model)
        for status, glob in (([D, I, D], [1, 1, 0]), ([I, D, I, D, D], [1, 0, 1, 1, 1])):
            xis = make_chain(oracle, d["xi_gt"], status, glob, seed=5)
            out[f"chain{len(status)}_{name}"] = (model, d["intr_gt"], d["board"], d["obs"], xis, status, glob)
    # failed projections: eta < 1e-3 and the alpha > 0.5 hemisphere test (eucm.h:46-54)
    d = sd.make_mono(sd.EUCM, 4, seed=930)
    xi = d["xi_gt"].copy(); xi[0, 2] -= 3.0; xi[1, 2] = -0.01
    out["sentinel_eucm"] = (sd.EUCM, d["intr_gt"], d["board"], d["obs"], [xi], [D], [0])
    intr = d["intr_gt"].copy(); intr[0] = 0.4
    out["sentinel_eucm_alpha04"] = (sd.EUCM, intr, d["board"], d["obs"], [xi], [D], [0])
    # small-angle branches (geometry_core.h:44,162; quaternion.h:34,88)
    xi = d["xi_gt"].copy(); xi[:, 3:] = 0; xi[:, :3] = [-0.4, -0.25, 0.8]
    xi[1, 3:] = [3e-6, -2e-6, 1e-6]; xi[2, 3:] = [9.9e-6, 0, 0]; xi[3, 3:] = [1.01e-5, 0, 0]
    out["small_angles"] = (sd.EUCM, d["intr_gt"], d["board"], d["obs"], [xi], [D], [0])
    # a 2x2 "IR" grid with z != 0 points (unified_calibration.cpp:234-250)
    d4 = sd.make_mono(sd.UCM, 3, seed=940, nx=2, ny=2, size=0.3)
    b = d4["board"].copy(); b[:, 2] = [0.0, 0.02, -0.01, 0.03]
    out["ir_grid_ucm"] = (sd.UCM, d4["intr_init"], b, d4["obs"], [d4["xi_init"]], [D], [0])
    return out


def main():
    ref, orc = Reference(), Oracle()
    lib = ref.lib
    blob = {}
    for name, (model, intr, board, obs, xis, status, glob) in cases(orc).items():
        o = ref.evaluate_batch(model, intr, board, obs, xis, status, glob, want_H=True)
        blob[f"{name}/model"] = np.array(model)
        blob[f"{name}/intr"] = intr; blob[f"{name}/board"] = board; blob[f"{name}/obs"] = obs
        blob[f"{name}/status"] = np.array(status); blob[f"{name}/is_global"] = np.array(glob)
        for e, x in enumerate(xis):
            blob[f"{name}/xi{e}"] = np.asarray(x)
        blob[f"{name}/r"] = o["r"]; blob[f"{name}/J_intr"] = o["J_intr"]; blob[f"{name}/H"] = o["H"]
        for e, j in enumerate(o["J_xi"]):
            blob[f"{name}/J_xi{e}"] = j
    # single-function probes
    u = sd.uniform(77, 1, 3 * 12).reshape(12, 3) * 2 - 1
    vecs = np.concatenate([u * 2.5, u[:4] * 1e-6, np.zeros((1, 3)), [[3.1, 0.2, -0.1]], [[1e-5, 0, 0]]])
    R = np.zeros((len(vecs), 9)); B = np.zeros((len(vecs), 9))
    for i, v in enumerate(vecs):
        v = np.ascontiguousarray(v)
        lib.vgref_rotation_matrix(dp(v), dp(R[i])); lib.vgref_inter_omega_rot(dp(v), dp(B[i]))
    blob["probe/rotvec"] = vecs; blob["probe/rotation_matrix"] = R; blob["probe/inter_omega_rot"] = B
    ta = np.concatenate([u[:6] * 0.5, u[6:12] * 2.0], axis=1)       # 6 transforms [t, r]
    tb = np.concatenate([u[6:12] * 0.3, u[:6] * 1.5], axis=1)
    comp = np.zeros((3, 6, 6))
    for k in range(3):
        for i in range(6):
            a, b = np.ascontiguousarray(ta[i]), np.ascontiguousarray(tb[i])
            lib.vgref_compose(dp(a), dp(b), dp(comp[k, i]), k)
    blob["probe/ta"] = ta; blob["probe/tb"] = tb; blob["probe/compose"] = comp
    lib.vgref_bound.restype = C.c_double
    for model, intr, name in ((sd.EUCM, sd.EUCM_GT_LEFT, "eucm"), (sd.UCM, sd.UCM_GT, "ucm"), (sd.MEI, sd.MEI_GT, "mei")):
        intr = np.ascontiguousarray(intr)
        K = len(intr)
        blob[f"probe/bounds_{name}"] = np.array([[lib.vgref_bound(model, dp(intr), i, 0), lib.vgref_bound(model, dp(intr), i, 1)]
                                                 for i in range(K)])
        uv = np.array([[700.0, 420.0], [100.0, 90.0], [1200.0, 750.0], [640.0, 400.0]])
        X = np.zeros((len(uv), 3)); ok = np.zeros(len(uv))
        for i in range(len(uv)):
            p = np.ascontiguousarray(uv[i])
            ok[i] = lib.vgref_reconstruct(model, dp(intr), dp(p), dp(X[i]))
        blob[f"probe/reconstruct_{name}_uv"] = uv; blob[f"probe/reconstruct_{name}_X"] = X; blob[f"probe/reconstruct_{name}_ok"] = ok
    path = os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(path, **blob)
    print(f"wrote {path}: {len(blob)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")
    priors(ref, orc)
    visual_cov(ref, orc)
    camera_api(ref)


def camera_points(model, seed, n=160):
    """3D points around the camera of a synthetic calibration: mostly in front, some beside and behind it (EUCM's
    eta < 1e-3 and hemisphere rejections, eucm.h:46-54), some on the optical axis."""
    u = sd.uniform(seed, 3, 3 * n).reshape(n, 3)
    X = np.stack([(2 * u[:, 0] - 1) * 1.5, (2 * u[:, 1] - 1) * 1.0, 0.2 + 1.3 * u[:, 2]], axis=1)
    X[n // 2: n // 2 + 20, 2] *= -1.0                     # behind the camera
    X[n // 2 + 20: n // 2 + 30, 2] = -0.3 * np.abs(X[n // 2 + 20: n // 2 + 30, 0])   # far off axis, slightly behind
    X[-4:, :2] = 0.0                                      # on the axis
    return np.ascontiguousarray(X)


def camera_api(ref):
    """ICamera::projectPoint / projectionJacobian / intrinsicJacobian / reconstructPoint of the reference's three
    camera classes on seeded point sets -> reference_camera.npz."""
    lib = ref.lib
    blob = {}
    cams = ((sd.EUCM, sd.EUCM_GT_LEFT, "eucm"), (sd.EUCM, np.array([0.4, 1.2, 300.0, 310.0, 640.0, 400.0]), "eucm_a04"),
            (sd.UCM, sd.UCM_GT, "ucm"), (sd.MEI, sd.MEI_GT, "mei"))
    for model, intr, name in cams:
        intr = np.ascontiguousarray(intr, dtype=np.float64)
        K = len(intr)
        X = camera_points(model, 3000 + K + len(name))
        n = len(X)
        uv = np.zeros((n, 2)); dpdx = np.zeros((n, 6)); dpda = np.zeros((n, 2 * K)); flags = np.zeros(n, dtype=np.int32)
        for i in range(n):
            flags[i] = lib.vgref_project_point(model, dp(intr), dp(X[i]), dp(uv[i]), dp(dpdx[i, :3]), dp(dpdx[i, 3:]),
                                               dp(dpda[i, :K]), dp(dpda[i, K:]))
        # back-projection of the projected points (and of a few raw pixels far from the centre)
        px = np.concatenate([uv[flags & 1 == 1][:40], [[5.0, 5.0], [1275.0, 795.0], [640.0, 400.0], [2500.0, -900.0]]])
        Xr = np.zeros((len(px), 3)); okr = np.zeros(len(px), dtype=np.int32)
        for i in range(len(px)):
            q = np.ascontiguousarray(px[i])
            okr[i] = lib.vgref_reconstruct(model, dp(intr), dp(q), dp(Xr[i]))
        blob[f"{name}/model"] = np.array(model); blob[f"{name}/intr"] = intr; blob[f"{name}/X"] = X
        blob[f"{name}/uv"] = uv; blob[f"{name}/dPdX"] = dpdx; blob[f"{name}/dPdintr"] = dpda; blob[f"{name}/flags"] = flags
        blob[f"{name}/px"] = px; blob[f"{name}/Xrec"] = Xr; blob[f"{name}/rec_ok"] = okr
    path = os.path.join(HERE, "reference_camera.npz")
    np.savez_compressed(path, **blob)
    print(f"wrote {path}: {len(blob)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


def visual_cov(ref, orc):
    """TrajectoryVisualQuality::visualCov (trajectory_generation.cpp:185-206) from the reference build, the object made by
    its own constructor -> reference_visualcov.npz."""
    blob = {}
    xi_board = np.array([0.3, -0.2, 0.1, 0.1, -0.2, 0.3])
    for model, name, intr in ((sd.EUCM, "eucm", sd.EUCM_GT_LEFT), (sd.UCM, "ucm", sd.UCM_GT), (sd.MEI, "mei", sd.MEI_GT)):
        d = sd.make_mono(model, 16, seed=70 + model)
        # camera poses such that camPose^-1 o xiBoard is a pose from which the whole board is visible
        cam = np.array([orc.compose(xi_board, x, "compose_inverse") for x in d["xi_gt"]])
        if model == sd.EUCM:      # a board seen edge-on, 36 of its 54 corners outside the model's domain: zero rows (eucm.h:141-150)
            cam[-1] = orc.compose(xi_board, [0.3, -0.25, 0.02, 0, 1.9, 0], "compose_inverse")
        blob[f"{name}/intr"] = np.asarray(intr); blob[f"{name}/cam_poses"] = cam
        blob[f"{name}/cov"] = ref.visual_cov(model, intr, xi_board, 9, 6, 0.1, 0.25, cam)
    blob["xi_board"] = xi_board; blob["board"] = sd.make_board(9, 6, 0.1); blob["feature_variance"] = np.array(0.25)
    path = os.path.join(HERE, "reference_visualcov.npz")
    np.savez_compressed(path, **blob)
    print(f"wrote {path}: {len(blob)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


def priors(ref, orc):
    """TransformationPrior / OdometryPrior (calib_cost_functions.h:64-108, .cpp:119-228) evaluated by the reference
    build on seeded inputs -> reference_priors.npz (a separate file: the vectors above stay byte-identical)."""
    n = 48
    u = sd.uniform(78, 1, 40 * n).reshape(n, 40) * 2 - 1
    blob = {}
    stiff = 0.5 + 50.0 * (u[:, 0:6] + 1)
    xp = np.concatenate([u[:, 6:9] * 0.5, u[:, 9:12] * 1.2], axis=1)
    xi = xp + u[:, 12:18] * 0.05
    xp[::8, 3:] *= 1e-6                          # small-angle branches of interOmegaRot / the quaternion
    xi[4::8] = xp[4::8]                          # zero error
    xi[5::8, 3:] += 2.5 * u[5::8, 18:21]         # large relative rotation (angle wrap in toRotationVector)
    r = np.zeros((n, 6)); J = np.zeros((n, 6, 6))
    for i in range(n):
        r[i], J[i] = ref.transformation_prior(stiff[i], xp[i], xi[i])
    blob.update({"tp/stiffness": stiff, "tp/xi_prior": xp, "tp/xi": xi, "tp/r": r, "tp/J": J})
    o1 = np.concatenate([u[:, 18:21], u[:, 21:24] * 0.7], axis=1)
    inc = np.concatenate([u[:, 24:27] * 0.1, u[:, 27:30] * 0.06], axis=1)
    inc[::6] *= 1e-3                             # below MIN_L / MIN_DELTA (.cpp:131-135)
    inc[3::6, 3:] = 0
    o2 = np.array([orc.compose(o1[i], inc[i]) for i in range(n)])
    x1 = o1 + u[:, 30:36] * 0.01
    x2 = o2 + u[:, 34:40] * 0.01
    errV, errW, lam = 0.1, 0.05, 0.02
    r = np.zeros((n, 6)); J1 = np.zeros((n, 6, 6)); J2 = np.zeros((n, 6, 6))
    for i in range(n):
        r[i], J1[i], J2[i] = ref.odometry_prior(errV, errW, lam, o1[i], o2[i], x1[i], x2[i])
    blob.update({"op/params": np.array([errV, errW, lam]), "op/odom1": o1, "op/odom2": o2, "op/xi1": x1, "op/xi2": x2,
                 "op/r": r, "op/J1": J1, "op/J2": J2})
    path = os.path.join(HERE, "reference_priors.npz")
    np.savez_compressed(path, **blob)
    print(f"wrote {path}: {len(blob)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


def odometry_cost(ref, orc):
    """OdometryCost (odometry_cost_function.cpp:144-267) evaluated by the reference build on seeded inputs ->
    reference_odometry_cost.npz: 40 blocks of 1 to 80 wheel-angle increments (straight runs, turns on the spot, increments
    below MIN_L / MIN_DELTA), the parameter block off the prior intrinsics, poses near and far from the odometry."""
    n = 40
    errV, errW, lam = 0.1, 0.05, 0.003
    intr_prior = np.array([0.1, 0.1, 0.5])
    intr = np.array([0.103, 0.0985, 0.52])
    blob = {"params": np.array([errV, errW, lam]), "intr_prior": intr_prior, "intr": intr}
    lens = [1 + int(v) for v in (sd.uniform(79, 1, n) * 80)]
    lens[0], lens[1] = 1, 80
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    u = sd.uniform(79, 2, 2 * off[-1]).reshape(-1, 2)
    dq = 0.05 + 0.5 * u
    for b in range(n):
        sl = slice(off[b], off[b + 1])
        if b % 5 == 1: dq[sl, 1] = dq[sl, 0]                       # straight
        if b % 5 == 2: dq[sl, 1] = -dq[sl, 0]                      # turning on the spot
        if b % 5 == 3: dq[sl] *= 1e-3                              # below MIN_L / MIN_DELTA
        if b % 5 == 4: dq[sl, 0] *= -1                             # reversing one wheel
    w = sd.uniform(79, 3, 12 * n).reshape(n, 12) * 2 - 1
    xi1 = np.concatenate([w[:, 0:3], w[:, 3:6] * 0.8], axis=1)
    xi2 = np.zeros((n, 6))
    r = np.zeros((n, 6)); J1 = np.zeros((n, 6, 6)); J2 = np.zeros((n, 6, 6)); J3 = np.zeros((n, 6, 3))
    zp = np.zeros((n, 6)); A = np.zeros((n, 6, 6))
    for b in range(n):
        q = dq[off[b]:off[b + 1]]
        zeta = ref.odometry_cost(errV, errW, lam, q, intr_prior, np.zeros(6), np.zeros(6), intr_prior)[4]      # the prior motion
        xi2[b] = orc.compose(xi1[b], zeta) + w[b, 6:12] * (0.02 if b % 3 else 0.3)
        r[b], J1[b], J2[b], J3[b], zp[b], A[b] = ref.odometry_cost(errV, errW, lam, q, intr_prior, xi1[b], xi2[b], intr)
    blob.update({"dq_offset": off, "dq": dq, "xi1": xi1, "xi2": xi2, "r": r, "J1": J1, "J2": J2, "J3": J3, "zeta_prior": zp, "A": A})
    path = os.path.join(HERE, "reference_odometry_cost.npz")
    np.savez_compressed(path, **blob)
    print(f"wrote {path}: {len(blob)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "odometry_cost":       # the other files stay byte-identical
        from oracle.pyoracle import Oracle, Reference
        odometry_cost(Reference(), Oracle())
    else:
        main()

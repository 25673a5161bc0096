"""CPU: the oracle's restatement of TransformationPrior / OdometryPrior (calib_cost_functions.h:64-108,
calib_cost_functions.cpp:119-228) against the committed vectors the REFERENCE build produced
(tests/golden/make_golden.py -> reference_priors.npz), and the oracle LM's two linear-algebra routes
(arrowhead elimination / dense factorisation) against each other."""
import os

import numpy as np
import pytest

import synthdata as sd
from oracle.pyoracle import OracleProblem
from util import assert_close

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_priors.npz"))
PIN_RTOL = 1e-13


def test_transformation_prior_matches_reference_vectors(oracle):
    for i in range(len(GOLD["tp/xi"])):
        r, J = oracle.transformation_prior(GOLD["tp/stiffness"][i], GOLD["tp/xi_prior"][i], GOLD["tp/xi"][i])
        assert_close(r, GOLD["tp/r"][i], f"tp[{i}]: r", PIN_RTOL)
        assert_close(J, GOLD["tp/J"][i], f"tp[{i}]: J", PIN_RTOL)


def test_odometry_prior_matches_reference_vectors(oracle):
    errV, errW, lam = GOLD["op/params"]
    for i in range(len(GOLD["op/xi1"])):
        r, J1, J2 = oracle.odometry_prior(errV, errW, lam, GOLD["op/odom1"][i], GOLD["op/odom2"][i],
                                          GOLD["op/xi1"][i], GOLD["op/xi2"][i])
        assert_close(r, GOLD["op/r"][i], f"op[{i}]: r", PIN_RTOL)
        assert_close(J1, GOLD["op/J1"][i], f"op[{i}]: J1", PIN_RTOL)
        assert_close(J2, GOLD["op/J2"][i], f"op[{i}]: J2", PIN_RTOL)


def test_prior_residual_is_zero_at_the_prior(oracle):
    xp = np.array([0.3, -0.2, 0.5, 0.4, -0.7, 0.2])
    r, J = oracle.transformation_prior([3, 4, 5, 6, 7, 8], xp, xp)
    assert np.abs(r).max() < 1e-14
    assert np.allclose(np.diag(J)[:3], [3, 4, 5])
    o1 = np.array([0.1, 0.2, 0.0, 0.0, 0.0, 0.3]); o2 = oracle.compose(o1, [0.05, 0, 0, 0, 0, 0.04])
    r, _, _ = oracle.odometry_prior(0.1, 0.1, 0.01, o1, o2, o1, o2)
    assert np.abs(r).max() < 1e-12


def test_odometry_jacobian_of_the_second_pose_is_a_derivative(oracle):
    """J2 = A blockdiag(R20, R20 M(r2)) is the derivative of the residual in the functor's own first-order model:
    check it against central differences of r wrt xi2 where the prior error is small (the functor's Jacobians
    are exact only at zero error, calib_cost_functions.cpp:190-213)."""
    o1 = np.array([0.2, -0.1, 0.0, 0.02, -0.01, 0.4]); o2 = oracle.compose(o1, [0.06, 0.01, 0, 0, 0, 0.05])
    _, _, J2 = oracle.odometry_prior(0.1, 0.1, 0.02, o1, o2, o1, o2)
    fd = np.zeros((6, 6)); h = 1e-6
    for k in range(6):
        e = np.zeros(6); e[k] = h
        rp, _, _ = oracle.odometry_prior(0.1, 0.1, 0.02, o1, o2, o1, o2 + e)
        rm, _, _ = oracle.odometry_prior(0.1, 0.1, 0.02, o1, o2, o1, o2 - e)
        fd[:, k] = (rp - rm) / (2 * h)
    assert np.abs(fd - J2).max() <= 1e-5 * np.abs(J2).max()


def test_dense_route_equals_arrowhead_route(oracle):
    """A constant, unobserved extra element switches the oracle LM to its dense factorisation: the solution must
    not change (the two routes solve the same normal equations)."""
    d = sd.make_mono(sd.EUCM, 12, seed=77)
    res = []
    for extra in (False, True):
        P = OracleProblem(oracle)
        cam = P.add_camera(sd.EUCM, d["intr_init"])
        xi = np.vstack([d["xi_init"], d["xi_init"][:1]]) if extra else d["xi_init"]
        tr = P.add_transform(xi, is_global=False)
        P.add_dataset(cam, d["board"], d["obs"], [tr], [0], seq_index=np.arange(12))
        if extra:
            P.set_pose_constant(tr, 12)
        s = P.solve()
        res.append((s.final_cost, P.camera(cam), P.transform(tr)[:12], s.iterations))
    assert abs(res[0][0] - res[1][0]) <= 1e-10 * res[0][0]
    assert np.abs(res[0][1] - res[1][1]).max() <= 1e-8 * np.abs(res[0][1]).max()
    assert np.abs(res[0][2] - res[1][2]).max() <= 1e-8


def test_odometry_problem_converges(oracle):
    d = sd.make_odometry(24)
    P = OracleProblem(oracle)
    cam = P.add_camera(sd.EUCM, d["intr_gt"], constant=True)
    bc = P.add_transform(d["xi_bc_init"], is_global=True)
    od = P.add_transform(d["xi_odom_init"], is_global=False)
    wB = P.add_transform(d["xi_wB_init"], is_global=True)
    P.add_dataset(cam, d["board"], d["obs"], [bc, od, wB], d["status"])
    assert P.add_odometry(od, d["err_v"], d["err_w"], d["lam"], d["odom"]) == 23
    P.set_pose_constant(od, 0)
    P.add_transformation_prior(bc, [10] * 6)
    o = oracle.default_options(); o.max_num_iterations = 12
    s = P.solve(o)
    assert s.final_cost < 1e-2 * s.initial_cost
    assert np.array_equal(P.transform(od)[0], d["xi_odom_init"][0])          # the anchor did not move
    assert np.abs(P.transform(od) - d["xi_odom_gt"]).max() < np.abs(d["odom"] - d["xi_odom_gt"]).max()


def test_loss_matches_ceres_corrector_restated_in_numpy(oracle):
    """SoftLOneLoss(a) + Ceres' Corrector on raw (r, J) in numpy against the oracle LM's cost, and the robustness it
    buys: gross outliers in two images pull the squared-loss intrinsics away, the robust ones stay."""
    d = sd.make_mono(sd.EUCM, 10, seed=5)
    obs = d["obs"].copy(); obs[3, :20] += 40.0; obs[7, 30:60] -= 25.0
    a = 2.0
    o = oracle.evaluate_batch(sd.EUCM, d["intr_init"], d["board"], obs, [d["xi_init"]], [0], [0])
    s = (o["r"] ** 2).sum(axis=1)
    rho = 2 * a * a * (np.sqrt(1 + s / (a * a)) - 1)
    P = OracleProblem(oracle)
    cam = P.add_camera(sd.EUCM, d["intr_init"]); tr = P.add_transform(d["xi_init"], is_global=False)
    ds = P.add_dataset(cam, d["board"], obs, [tr], [0])
    P.set_loss(ds, a)
    assert abs(P.evaluate() - 0.5 * rho.sum()) <= 1e-12 * rho.sum()
    res = {}
    for loss in (0.0, 1.0):
        P = OracleProblem(oracle)
        cam = P.add_camera(sd.EUCM, d["intr_init"]); tr = P.add_transform(d["xi_init"], is_global=False)
        ds = P.add_dataset(cam, d["board"], obs, [tr], [0])
        P.set_loss(ds, loss)
        P.solve()
        res[loss] = np.abs(P.camera(cam)[2:] - d["intr_gt"][2:]).max()
    assert res[1.0] < 0.5 * res[0.0]


VCOV = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_visualcov.npz"))


@pytest.mark.parametrize("model,name", [(sd.EUCM, "eucm"), (sd.UCM, "ucm"), (sd.MEI, "mei")])
def test_visual_cov_matches_reference_vectors(oracle, model, name):
    """TrajectoryVisualQuality::visualCov (trajectory_generation.cpp:185-206): the oracle against the
    reference build's own object (constructed through its constructor from an in-memory property tree)."""
    got = oracle.visual_cov(model, VCOV[f"{name}/intr"], VCOV["xi_board"], VCOV["board"], float(VCOV["feature_variance"]),
                            VCOV[f"{name}/cam_poses"])
    want = VCOV[f"{name}/cov"]
    for k in range(len(want)):
        assert np.abs(got[k] - want[k]).max() <= 1e-11 * np.abs(want[k]).max()
    # the covariance of a localisation is symmetric positive definite
    w = np.linalg.eigvalsh((want[0] + want[0].T) / 2)
    assert w[0] > 0


# ---- OdometryCost (src/calibration/odometry_cost_function.cpp; the residual block of "odometry_intrinsic" datasets) ----
OC = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_odometry_cost.npz"))


def _oc_block(b):
    off = OC["dq_offset"]
    return OC["dq"][off[b]:off[b + 1]]


def test_odometry_cost_matches_reference_vectors(oracle):
    errV, errW, lam = OC["params"]
    for b in range(len(OC["r"])):
        r, J1, J2, J3, zp, A = oracle.odometry_cost(errV, errW, lam, _oc_block(b), OC["intr_prior"], OC["xi1"][b], OC["xi2"][b], OC["intr"])
        for got, name in ((r, "r"), (J1, "J1"), (J2, "J2"), (J3, "J3"), (zp, "zeta_prior"), (A, "A")):
            want = OC[name][b]
            assert np.abs(got - want).max() <= 1e-13 * max(1.0, np.abs(want).max()), (b, name)


def test_odometry_cost_jacobians_are_derivatives(oracle):
    """Central differences of the residual at the odometry (delta = 0).  d r / d intrinsics is exact for a single
    increment; for longer chains the reference's calc_acc is itself an approximation (it differentiates the planar
    position through the first-order term only: a few per cent off at 80 increments) -- the restatement follows the
    reference, so the longer chains are held loosely.  The translation columns of d r / d xi2 are exact."""
    errV, errW, lam = OC["params"]
    for b in (0, 1, 7, 12, 26):
        q, ip = _oc_block(b), OC["intr_prior"]
        zeta = oracle.odometry_cost(errV, errW, lam, q, ip, np.zeros(6), np.zeros(6), OC["intr"])[4]
        f = lambda x1, x2, it: oracle.odometry_cost(errV, errW, lam, q, ip, x1, x2, it)[0]
        x1 = OC["xi1"][b]
        x2 = oracle.compose(x1, oracle.odometry_cost(errV, errW, lam, q, OC["intr"], np.zeros(6), np.zeros(6), OC["intr"])[4])
        _, J1, J2, J3, _, A = oracle.odometry_cost(errV, errW, lam, q, ip, x1, x2, OC["intr"])
        h = 1e-6
        num3 = np.stack([(f(x1, x2, OC["intr"] + h * e) - f(x1, x2, OC["intr"] - h * e)) / (2 * h) for e in np.eye(3)], axis=1)
        assert np.abs(num3 - J3).max() <= (2e-6 if len(q) == 1 else 0.1) * max(1.0, np.abs(J3).max()), b
        num2 = np.stack([(f(x1, x2 + h * e, OC["intr"]) - f(x1, x2 - h * e, OC["intr"])) / (2 * h) for e in np.eye(6)], axis=1)
        assert np.abs(num2[:, :3] - J2[:, :3]).max() <= 2e-6 * max(1.0, np.abs(J2).max()), b      # translation columns
        assert zeta.shape == (6,) and A.shape == (6, 6)


def test_odometry_cost_equals_odometry_prior_on_the_same_motion(oracle):
    """With the parameter block at the prior intrinsics the integrated motion IS the prior: residual and the first two
    Jacobian blocks are OdometryPrior's for odom1 = 0, odom2 = that motion."""
    errV, errW, lam = OC["params"]
    for b in (2, 9, 33):
        q, ip = _oc_block(b), OC["intr_prior"]
        r, J1, J2, _, zp, _ = oracle.odometry_cost(errV, errW, lam, q, ip, OC["xi1"][b], OC["xi2"][b], ip)
        r0, K1, K2 = oracle.odometry_prior(errV, errW, lam, np.zeros(6), zp, OC["xi1"][b], OC["xi2"][b])
        assert np.abs(r - r0).max() < 1e-12 and np.abs(J1 - K1).max() < 1e-12 and np.abs(J2 - K2).max() < 1e-12

"""GPU parity of the 6-residual blocks that share the global problem with the reprojection blocks:
TransformationPrior and OdometryPrior (calib_cost_functions.h:64-108, calib_cost_functions.cpp:119-228), at the
functor level against the REFERENCE build's vectors and the oracle, and at the problem level (cost, LM trajectory,
final parameters) against the oracle LM, whose coupled-pose step is a dense factorisation -- a different route to
the same normal equations than the CUDA engine's block-tridiagonal elimination.

Tolerances: functor outputs 1e-11 of each block's scale; LM 1e-8 relative on costs / parameters after a fixed
number of iterations (north_star allows 1e-6)."""
import os

import numpy as np
import pytest

import synthdata as sd
from oracle.pyoracle import OracleProblem
from util import assert_close

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_priors.npz"))
FUNCTOR_RTOL = 1e-11
LM_RTOL = 1e-8


def test_transformation_prior_matches_reference_vectors(gpu):
    r, J = gpu.eval_transformation_prior(GOLD["tp/stiffness"], GOLD["tp/xi_prior"], GOLD["tp/xi"])
    for i in range(len(r)):
        assert_close(r[i], GOLD["tp/r"][i], f"tp[{i}]: r", FUNCTOR_RTOL)
        assert_close(J[i], GOLD["tp/J"][i], f"tp[{i}]: J", FUNCTOR_RTOL)
    r2, none = gpu.eval_transformation_prior(GOLD["tp/stiffness"], GOLD["tp/xi_prior"], GOLD["tp/xi"], want_J=False)
    assert none is None and np.array_equal(r, r2)


def test_odometry_prior_matches_reference_vectors(gpu):
    errV, errW, lam = GOLD["op/params"]
    r, J1, J2 = gpu.eval_odometry_prior(errV, errW, lam, GOLD["op/odom1"], GOLD["op/odom2"], GOLD["op/xi1"], GOLD["op/xi2"])
    for i in range(len(r)):
        assert_close(r[i], GOLD["op/r"][i], f"op[{i}]: r", FUNCTOR_RTOL)
        assert_close(J1[i], GOLD["op/J1"][i], f"op[{i}]: J1", FUNCTOR_RTOL)
        assert_close(J2[i], GOLD["op/J2"][i], f"op[{i}]: J2", FUNCTOR_RTOL)


def test_prior_functors_match_oracle_on_random_inputs(gpu, oracle):
    n = 500
    u = sd.uniform(4711, 1, 36 * n).reshape(n, 36) * 2 - 1
    stiff = 1 + 40 * (u[:, :6] + 1); xp = u[:, 6:12] * [0.5, 0.5, 0.5, 1.5, 1.5, 1.5]; xi = xp + 0.1 * u[:, 12:18]
    r, J = gpu.eval_transformation_prior(stiff, xp, xi)
    for i in range(0, n, 7):
        ro, Jo = oracle.transformation_prior(stiff[i], xp[i], xi[i])
        assert_close(r[i], ro, f"tp[{i}]: r", FUNCTOR_RTOL); assert_close(J[i], Jo, f"tp[{i}]: J", FUNCTOR_RTOL)
    o1 = u[:, 18:24] * [1, 1, 1, 0.8, 0.8, 0.8]
    o2 = np.array([oracle.compose(o1[i], u[i, 24:30] * [0.2, 0.2, 0.2, 0.1, 0.1, 0.1]) for i in range(n)])
    x1 = o1 + 0.02 * u[:, 30:36]; x2 = o2 - 0.02 * u[:, 28:34]
    r, J1, J2 = gpu.eval_odometry_prior(0.07, 0.03, 0.015, o1, o2, x1, x2)
    for i in range(0, n, 7):
        ro, J1o, J2o = oracle.odometry_prior(0.07, 0.03, 0.015, o1[i], o2[i], x1[i], x2[i])
        assert_close(r[i], ro, f"op[{i}]: r", FUNCTOR_RTOL)
        assert_close(J1[i], J1o, f"op[{i}]: J1", FUNCTOR_RTOL); assert_close(J2[i], J2o, f"op[{i}]: J2", FUNCTOR_RTOL)
    assert gpu.eval_transformation_prior(np.zeros((0, 6)), np.zeros((0, 6)), np.zeros((0, 6)))[0].shape == (0, 6)


def build_odometry(P, d, cam_const=True, odo=True, anchor=True, prior=True, board_const=False):
    cam = P.add_camera(sd.EUCM, d["intr_gt"] if cam_const else d["intr_init"], constant=cam_const)
    bc = P.add_transform(d["xi_bc_init"], is_global=True)
    od = P.add_transform(d["xi_odom_init"], is_global=False)
    wB = P.add_transform(d["xi_wB_gt"] if board_const else d["xi_wB_init"], is_global=True, constant=board_const)
    P.add_dataset(cam, d["board"], d["obs"], [bc, od, wB], d["status"])
    if odo:
        P.add_odometry(od, d["err_v"], d["err_w"], d["lam"], d["odom"])
    if anchor:
        P.set_pose_constant(od, 0)
    if prior:
        P.add_transformation_prior(bc, [10, 10, 10, 20, 20, 20])
    return cam, bc, od, wB


def rel(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-2)))


def compare_solutions(G, O, gids, oids, iters, oracle, tol=LM_RTOL, same_steps=True):
    """Both LMs run `iters` iterations from the same start.  same_steps: the accept / reject sequence must agree
    too (only meaningful while the iteration is still away from convergence -- at the minimum the cost change of a
    trial step is rounding noise and so is its acceptance)."""
    og = G.default_options(); og.max_num_iterations = iters
    oo = oracle.default_options(); oo.max_num_iterations = iters
    sg, so = G.solve(og), O.solve(oo)
    assert abs(sg.initial_cost - so.initial_cost) <= (1e-10 if same_steps else tol) * so.initial_cost
    if same_steps:
        assert (sg.iterations, sg.num_successful, sg.num_unsuccessful) == (so.iterations, so.num_successful, so.num_unsuccessful)
    assert abs(sg.final_cost - so.final_cost) <= tol * so.final_cost
    assert rel(G.camera(gids[0]), O.camera(oids[0])) < tol
    for g, o in zip(gids[1:], oids[1:]):
        assert rel(G.transform(g), O.transform(o)) < tol
    return sg, so


@pytest.mark.parametrize("cam_const", [True, False])
def test_odometry_problem_matches_oracle_lm(gpu, oracle, cam_const):
    """Odometry blocks couple consecutive poses (block-tridiagonal pose part), the first pose is anchored, the
    camera extrinsic carries a TransformationPrior: cost, accept/reject sequence and parameters after 8 iterations."""
    d = sd.make_odometry(40)
    G, O = gpu.Problem(), OracleProblem(oracle)
    gids, oids = build_odometry(G, d, cam_const), build_odometry(O, d, cam_const)
    assert abs(G.evaluate() - O.evaluate()) <= 1e-11 * O.evaluate()
    sg, so = compare_solutions(G, O, gids, oids, 8, oracle)
    assert sg.final_cost < 1e-2 * sg.initial_cost
    assert np.array_equal(G.transform(gids[2])[0], d["xi_odom_init"][0])     # the anchor did not move
    # the odometry pulled the poses towards the truth
    assert np.abs(G.transform(gids[2]) - d["xi_odom_gt"]).max() < np.abs(d["odom"] - d["xi_odom_gt"]).max()


def test_long_chain_and_wide_shared_block(gpu, oracle):
    """150 coupled poses, two free globals + a free camera (18 shared columns)."""
    d = sd.make_odometry(150, seed=777)
    G, O = gpu.Problem(), OracleProblem(oracle)
    gids, oids = build_odometry(G, d, False), build_odometry(O, d, False)
    compare_solutions(G, O, gids, oids, 6, oracle)


def test_odometry_without_anchor_or_prior(gpu, oracle):
    """No constant element, constant board: the chain alone (every pose free and coupled)."""
    d = sd.make_odometry(25, seed=99)
    G, O = gpu.Problem(), OracleProblem(oracle)
    gids = build_odometry(G, d, True, anchor=False, prior=False, board_const=True)
    oids = build_odometry(O, d, True, anchor=False, prior=False, board_const=True)
    compare_solutions(G, O, gids, oids, 6, oracle)


def test_constant_elements_without_odometry(gpu, oracle):
    """SetParameterBlockConstant on single elements of a sequence in a plain grid problem."""
    d = sd.make_mono(sd.EUCM, 16, seed=321)
    G, O = gpu.Problem(), OracleProblem(oracle)
    ids = []
    for P in (G, O):
        cam = P.add_camera(sd.EUCM, d["intr_init"])
        tr = P.add_transform(d["xi_init"], is_global=False)
        P.add_dataset(cam, d["board"], d["obs"], [tr], [0])
        for k in (0, 5, 15):
            P.set_pose_constant(tr, k)
        ids.append((cam, tr))
    compare_solutions(G, O, ids[0], ids[1], 5, oracle)
    # on to convergence: three boards frozen at perturbed poses leave alpha / beta in a flat valley, where the two
    # LMs stop 2e-7 apart (north_star's bound is 1e-6); the objective itself is held to 1e-8 of its value
    sg, so = compare_solutions(G, O, ids[0], ids[1], 10, oracle, tol=1e-6, same_steps=False)
    assert abs(sg.final_cost - so.final_cost) <= LM_RTOL * so.final_cost
    for k in (0, 5, 15):
        assert np.array_equal(G.transform(ids[0][1])[k], d["xi_init"][k])


def test_priors_on_sequence_elements_and_globals(gpu, oracle):
    """TransformationPrior on two elements of a sequence (independent poses that gather a prior record) and on the
    stereo extrinsic (shared block), with an explicit prior value."""
    s = sd.make_stereo(12, seed=555)
    G, O = gpu.Problem(), OracleProblem(oracle)
    ids = []
    for P in (G, O):
        c1 = P.add_camera(sd.EUCM, s["intr1_init"]); c2 = P.add_camera(sd.EUCM, s["intr2_init"])
        t12 = P.add_transform(s["xi12_init"], is_global=True)
        tb = P.add_transform(s["xi_init"], is_global=False)
        P.add_dataset(c1, s["board"], s["obs1"], [tb], [0])
        P.add_dataset(c2, s["board"], s["obs2"], [t12, tb], [1, 0])
        P.add_transformation_prior(t12, [50, 50, 50, 100, 100, 100])                      # prior = current value
        P.add_transformation_prior(tb, [5, 5, 5, 9, 9, 9], index=3, xi_prior=s["xi_gt"][3])
        P.add_transformation_prior(tb, [2, 3, 4, 5, 6, 7], index=7)
        ids.append((c1, t12, tb))
    assert abs(G.evaluate() - O.evaluate()) <= 1e-11 * O.evaluate()
    compare_solutions(G, O, ids[0], ids[1], 5, oracle)
    compare_solutions(G, O, ids[0], ids[1], 6, oracle, same_steps=False)       # on to convergence
    assert rel(G.camera(1), O.camera(1)) < LM_RTOL


def test_mixed_problem_free_poses_and_a_chain(gpu, oracle):
    """One dataset with independent board poses and one with an odometry chain share the camera."""
    d = sd.make_odometry(30, seed=4242)
    m = sd.make_mono(sd.EUCM, 30, seed=11, intr_guess=d["intr_init"])
    G, O = gpu.Problem(), OracleProblem(oracle)
    ids = []
    for P in (G, O):
        cam = P.add_camera(sd.EUCM, d["intr_init"])
        tr = P.add_transform(m["xi_init"], is_global=False)
        P.add_dataset(cam, m["board"], m["obs"], [tr], [0])
        bc = P.add_transform(d["xi_bc_init"], is_global=True)
        od = P.add_transform(d["xi_odom_init"], is_global=False)
        wB = P.add_transform(d["xi_wB_init"], is_global=True)
        P.add_dataset(cam, d["board"], d["obs"], [bc, od, wB], d["status"])
        P.add_odometry(od, d["err_v"], d["err_w"], d["lam"], d["odom"])
        P.set_pose_constant(od, 0)
        ids.append((cam, tr, bc, od, wB))
    compare_solutions(G, O, ids[0], ids[1], 8, oracle)


def test_prior_api_errors(gpu):
    d = sd.make_odometry(5)
    P = gpu.Problem()
    g = P.add_transform(d["xi_bc_init"], is_global=True)
    sq = P.add_transform(d["xi_odom_init"], is_global=False)
    with pytest.raises(gpu.VisgeomError, match="Odometry must be a sequence"):
        P.add_odometry(g, 0.1, 0.1, 0.01, d["odom"][:1])
    with pytest.raises(gpu.VisgeomError, match="one odometry reading per sequence element"):
        P.add_odometry(sq, 0.1, 0.1, 0.01, d["odom"][:3])
    with pytest.raises(gpu.VisgeomError, match="out of range"):
        P.add_transformation_prior(sq, [1] * 6, index=9)
    with pytest.raises(gpu.VisgeomError, match="not a sequence"):
        P.set_pose_constant(g, 0)
    P.add_odometry(sq, 0.1, 0.1, 0.01, d["odom"])
    with pytest.raises(gpu.VisgeomError, match="already has odometry"):
        P.add_odometry(sq, 0.1, 0.1, 0.01, d["odom"])


VCOV = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_visualcov.npz"))


@pytest.mark.parametrize("model,name", [(sd.EUCM, "eucm"), (sd.UCM, "ucm"), (sd.MEI, "mei")])
def test_visual_cov_matches_reference_and_oracle(gpu, oracle, model, name):
    """TrajectoryVisualQuality::visualCov (trajectory_generation.cpp:185-206) through vg_visual_cov: the
    reference build's vectors (a pose with 36 of 54 corners outside the model's domain included), then 300 random
    poses against the oracle.  Tolerance 1e-9 of each matrix's largest entry (the 6 x 6 inverse amplifies rounding
    by the condition number of J^T J, ~1e3 here)."""
    fv = float(VCOV["feature_variance"])
    got = gpu.visual_cov(model, VCOV[f"{name}/intr"], VCOV["xi_board"], VCOV["board"], fv, VCOV[f"{name}/cam_poses"])
    want = VCOV[f"{name}/cov"]
    for k in range(len(want)):
        assert np.isfinite(got[k]).all()
        assert np.abs(got[k] - want[k]).max() <= 1e-9 * np.abs(want[k]).max(), (name, k)
    d = sd.make_mono(model, 300, seed=4100 + model)
    xi_board = np.array([-0.2, 0.4, 0.3, 0.3, 0.1, -0.4])
    cam = np.array([oracle.compose(xi_board, x, "compose_inverse") for x in d["xi_gt"]])
    got = gpu.visual_cov(model, d["intr_gt"], xi_board, d["board"], 0.04, cam)
    want = oracle.visual_cov(model, d["intr_gt"], xi_board, d["board"], 0.04, cam)
    rel_err = np.abs(got - want).max(axis=(1, 2)) / np.abs(want).max(axis=(1, 2))
    assert rel_err.max() <= 1e-9, rel_err.max()
    assert gpu.visual_cov(model, d["intr_gt"], xi_board, d["board"], 0.04, np.zeros((0, 6))).shape == (0, 6, 6)
    with pytest.raises(gpu.VisgeomError, match="bad arguments"):
        gpu.visual_cov(model, d["intr_gt"], xi_board, d["board"], -1.0, cam)


# ---- OdometryCost (src/calibration/odometry_cost_function.cpp) at the functor level -------------------------------------
OC = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_odometry_cost.npz"))


def test_odometry_cost_matches_reference_vectors(gpu):
    errV, errW, lam = OC["params"]
    off = OC["dq_offset"]
    blocks = [OC["dq"][off[b]:off[b + 1]] for b in range(len(OC["r"]))]
    r, J1, J2, J3 = gpu.eval_odometry_cost(errV, errW, lam, blocks, OC["intr_prior"], OC["xi1"], OC["xi2"], OC["intr"])
    for b in range(len(r)):
        assert_close(r[b], OC["r"][b], f"oc[{b}]: r", FUNCTOR_RTOL)
        assert_close(J1[b], OC["J1"][b], f"oc[{b}]: J1", FUNCTOR_RTOL)
        assert_close(J2[b], OC["J2"][b], f"oc[{b}]: J2", FUNCTOR_RTOL)
        assert_close(J3[b], OC["J3"][b], f"oc[{b}]: J3", FUNCTOR_RTOL)
    r2, a, b_, c = gpu.eval_odometry_cost(errV, errW, lam, blocks, OC["intr_prior"], OC["xi1"], OC["xi2"], OC["intr"], want_J=False)
    assert a is None and b_ is None and c is None and np.array_equal(r, r2)


def test_odometry_cost_matches_oracle_on_random_inputs(gpu, oracle):
    n = 300
    rng = np.random.default_rng(815)
    blocks = [rng.uniform(-0.4, 0.6, (int(rng.integers(1, 120)), 2)) * rng.choice([1.0, 1.0, 1e-3]) for _ in range(n)]
    ip = np.array([0.08, 0.081, 0.43]); it = ip * (1 + rng.normal(0, 0.03, 3))
    xi1 = np.concatenate([rng.normal(0, 1, (n, 3)), rng.normal(0, 0.6, (n, 3))], axis=1)
    xi2 = xi1 + np.concatenate([rng.normal(0, 0.4, (n, 3)), rng.normal(0, 0.3, (n, 3))], axis=1)
    r, J1, J2, J3 = gpu.eval_odometry_cost(0.07, 0.09, 0.004, blocks, ip, xi1, xi2, it)
    for b in range(n):
        o = oracle.odometry_cost(0.07, 0.09, 0.004, blocks[b], ip, xi1[b], xi2[b], it)
        for got, want, name in ((r[b], o[0], "r"), (J1[b], o[1], "J1"), (J2[b], o[2], "J2"), (J3[b], o[3], "J3")):
            assert_close(got, want, f"oc[{b}]: {name}", 1e-10)


def test_odometry_cost_api_errors(gpu):
    import ctypes as C
    z6 = np.zeros((1, 6)); i3 = np.array([0.1, 0.1, 0.5])
    with pytest.raises(gpu.VisgeomError):
        gpu.eval_odometry_cost(0.1, 0.1, 0.01, [np.zeros((0, 2))], i3, z6, z6, i3)      # a block without increments
    with pytest.raises(gpu.VisgeomError):
        gpu.eval_odometry_cost(0.1, 0.1, 0.0, [np.ones((3, 2))], i3, z6, z6, i3)        # lambda = 0 divides by zero
    r, *_ = gpu.eval_odometry_cost(0.1, 0.1, 0.01, [], i3, np.zeros((0, 6)), np.zeros((0, 6)), i3)
    assert r.shape == (0, 6)

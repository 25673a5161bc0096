"""CPU: independent checks of the oracle itself -- finite differences of every Jacobian block
(including INVERSE chain elements), sentinel behaviour, and the oracle LM against
scipy.optimize.least_squares on the same objective."""
import numpy as np
import pytest

import synthdata as sd
from oracle.pyoracle import OracleProblem
from util import assert_close, assert_close_hessian

D, I = 0, 1


def fd_block(oracle, args, which, e=None, eps=1e-6):
    model, intr, board, obs, xis, status, glob = args
    base = oracle.evaluate_batch(*args)

    def res(intr_, xis_):
        return oracle.evaluate_batch(model, intr_, board, obs, xis_, status, glob, want_J=False)["r"]
    if which == "intr":
        J = base["J_intr"]
        for k in range(len(intr)):
            h = eps * max(1.0, abs(intr[k]))
            p, m = intr.copy(), intr.copy(); p[k] += h; m[k] -= h
            fd = (res(p, xis) - res(m, xis)) / (2 * h)
            assert np.abs(fd - J[:, :, k]).max() <= 2e-6 * max(1.0, np.abs(J[:, :, k]).max()), ("intr", k)
    else:
        J = base["J_xi"][e]
        for k in range(6):
            xp = [x.copy() for x in xis]; xm = [x.copy() for x in xis]
            xp[e][..., k] += eps; xm[e][..., k] -= eps
            fd = (res(intr, xp) - res(intr, xm)) / (2 * eps)
            assert np.abs(fd - J[:, :, k]).max() <= 2e-6 * max(1.0, np.abs(J[:, :, k]).max()), ("xi", e, k)


@pytest.mark.parametrize("model", [sd.EUCM, sd.UCM, sd.MEI])
def test_jacobians_are_true_derivatives_mono(oracle, model):
    d = sd.make_mono(model, 6, seed=50 + model)
    args = (model, d["intr_gt"], d["board"], d["obs"], [d["xi_gt"]], [D], [0])
    fd_block(oracle, args, "intr")
    fd_block(oracle, args, "xi", 0)


def test_jacobians_are_true_derivatives_inverse_chain(oracle):
    s = sd.make_stereo(5, seed=60)
    args = (sd.EUCM, s["intr2_gt"], s["board"], s["obs2"], [s["xi12_gt"], s["xi_gt"]], [I, D], [1, 0])
    fd_block(oracle, args, "intr")
    fd_block(oracle, args, "xi", 0)
    fd_block(oracle, args, "xi", 1)


def test_sentinel_rows(oracle):
    d = sd.make_mono(sd.EUCM, 4, seed=70)
    xi = d["xi_gt"].copy(); xi[:, 2] -= 5.0
    o = oracle.evaluate_batch(sd.EUCM, d["intr_gt"], d["board"], d["obs"], [xi], [D], [0], want_H=True)
    assert (o["r"] == 1e15).all() and (o["J_intr"] == 0).all() and (o["J_xi"][0] == 0).all()
    assert np.allclose(o["H"][:, -1], 2 * d["P"] * 1e30)


def test_oracle_lm_matches_scipy(oracle):
    """The converged minimum is a property of the objective: the oracle LM (Ceres-style trust region,
    arrowhead Schur) and scipy's trust-region-reflective solver must agree on it."""
    from scipy.optimize import least_squares
    d = sd.make_mono(sd.EUCM, 8, seed=20241)
    n, K = d["n_img"], d["K"]
    lo = np.array([oracle.lib.vgo_lower_bound(sd.EUCM, i) for i in range(K)])
    hi = np.array([oracle.lib.vgo_upper_bound(sd.EUCM, i) for i in range(K)])

    def unpack(x):
        return x[:K], x[K:].reshape(n, 6)

    def fun(x):
        intr, xi = unpack(x)
        return oracle.evaluate_batch(sd.EUCM, intr, d["board"], d["obs"], [xi], [D], [0], want_J=False)["r"].ravel()

    def jac(x):
        intr, xi = unpack(x)
        o = oracle.evaluate_batch(sd.EUCM, intr, d["board"], d["obs"], [xi], [D], [0])
        J = np.zeros((n * 2 * d["P"], K + 6 * n))
        J[:, :K] = o["J_intr"].reshape(-1, K)
        for i in range(n):
            J[i * 2 * d["P"]:(i + 1) * 2 * d["P"], K + 6 * i:K + 6 * i + 6] = o["J_xi"][0][i]
        return J
    x0 = np.concatenate([d["intr_init"], d["xi_init"].ravel()])
    lb = np.concatenate([lo, np.full(6 * n, -np.inf)]); ub = np.concatenate([hi, np.full(6 * n, np.inf)])
    sp = least_squares(fun, x0, jac=jac, bounds=(lb, ub), method="trf", xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=200)
    P = OracleProblem(oracle)
    cam = P.add_camera(sd.EUCM, d["intr_init"]); tr = P.add_transform(d["xi_init"], False)
    P.add_dataset(cam, d["board"], d["obs"], [tr], [D])
    s = P.solve()
    assert abs(s.final_cost - sp.cost) <= 1e-9 * sp.cost
    assert np.max(np.abs(P.camera(cam) - sp.x[:K]) / np.abs(sp.x[:K])) < 1e-6
    assert np.abs(P.transform(tr).ravel() - sp.x[K:]).max() < 1e-6


def test_synthetic_data_is_deterministic():
    a, b = sd.make_mono(sd.EUCM, 5, seed=1), sd.make_mono(sd.EUCM, 5, seed=1)
    assert (a["obs"] == b["obs"]).all() and (a["xi_init"] == b["xi_init"]).all()
    # frozen values: any change to the generator invalidates BENCH comparisons across rounds
    d = sd.make_mono(sd.EUCM, 20, seed=20241)
    assert d["board"].shape == (54, 3) and tuple(d["board"][10]) == (0.1, 0.1, 0.0)     # unified_calibration.cpp:286-292
    assert abs(float(d["obs"].sum()) - 1093379.0) < 1e6   # sanity, not a fingerprint
    assert np.isfinite(d["obs"]).all() and (d["obs"] > 0).all() and (d["obs"][:, 0::2] < 1280).all()


@pytest.mark.parametrize("model", [sd.EUCM, sd.UCM, sd.MEI])
def test_tuned_cpu_variant_equals_the_faithful_oracle(oracle, model):
    """oracle_tuned.c (bench.py's second CPU baseline) produces the faithful restatement's outputs."""
    from oracle.pyoracle import TunedOracle
    from test_eval_gpu import make_chain
    tuned = TunedOracle()
    d = sd.make_mono(model, 40, seed=60 + model)
    xi_fail = d["xi_init"].copy(); xi_fail[3, 2] = -0.05                     # behind the camera: sentinel rows
    cases = [(model, d["intr_init"], d["board"], d["obs"], [xi_fail], [0], [0])]
    st, gl = [1, 0, 1, 0, 0], [1, 0, 1, 1, 1]
    cases.append((model, d["intr_gt"], d["board"], d["obs"], make_chain(oracle, d["xi_gt"], st, gl, seed=5), st, gl))
    for args in cases:
        a, b = oracle.evaluate_batch(*args, want_H=True), tuned.evaluate_batch(*args, want_H=True, threads=2)
        assert ((a["r"] == 1e15) == (b["r"] == 1e15)).all()
        assert_close(b["r"], a["r"], "r", 1e-12)
        assert_close(b["J_intr"], a["J_intr"], "J_intr", 1e-12)
        for x, y in zip(b["J_xi"], a["J_xi"]):
            assert_close(x, y, "J_xi", 1e-12)
        assert_close_hessian(b["H"], a["H"], "H", 1e-11)

"""CPU: the oracle (oracle/visgeom_oracle.c) against the committed golden vectors produced by the
REFERENCE's own code (tests/golden/make_golden.py -> oracle/_ref), and -- where the reference tree
is present -- against the reference build live on larger random cases.  This is what pins the oracle."""
import os

import numpy as np
import pytest

import synthdata as sd
from util import assert_close, assert_close_hessian

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))
CASES = sorted({k.split("/")[0] for k in GOLD.files if not k.startswith("probe/")})
PIN_RTOL = 1e-13     # oracle vs reference code: same algorithm, same operation order


def load_case(name):
    g = lambda k: GOLD[f"{name}/{k}"]
    L = len(g("status"))
    xis = [g(f"xi{e}") for e in range(L)]
    return (int(g("model")), g("intr"), g("board"), g("obs"), xis, list(g("status")), list(g("is_global"))), \
        dict(r=g("r"), J_intr=g("J_intr"), H=g("H"), J_xi=[g(f"J_xi{e}") for e in range(L)])


def test_golden_file_covers_the_path():
    assert len(CASES) >= 15
    assert any(c.startswith("chain5") for c in CASES) and "stereo_cam2" in CASES and "sentinel_eucm" in CASES


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_vectors(oracle, name):
    args, want = load_case(name)
    got = oracle.evaluate_batch(*args, want_H=True)
    assert_close(got["r"], want["r"], f"{name}: r", PIN_RTOL)
    assert ((got["r"] == 1e15) == (want["r"] == 1e15)).all()
    assert_close(got["J_intr"], want["J_intr"], f"{name}: J_intr", PIN_RTOL)
    for e, (a, b) in enumerate(zip(got["J_xi"], want["J_xi"])):
        assert_close(a, b, f"{name}: J_xi[{e}]", PIN_RTOL)
    assert_close_hessian(got["H"], want["H"], f"{name}: H", 1e-12)


def test_geometry_probes(oracle):
    v = GOLD["probe/rotvec"]
    for i in range(len(v)):
        assert np.abs(oracle.rotation_matrix(v[i]).ravel() - GOLD["probe/rotation_matrix"][i]).max() < 1e-15
        assert np.abs(oracle.inter_omega_rot(v[i]).ravel() - GOLD["probe/inter_omega_rot"][i]).max() < 1e-15
    ta, tb, comp = GOLD["probe/ta"], GOLD["probe/tb"], GOLD["probe/compose"]
    for k, kind in enumerate(("compose", "compose_inverse", "inverse_compose")):
        for i in range(len(ta)):
            assert np.abs(oracle.compose(ta[i], tb[i], kind) - comp[k, i]).max() < 1e-14, (kind, i)


@pytest.mark.parametrize("model,name,intr", [(sd.EUCM, "eucm", sd.EUCM_GT_LEFT), (sd.UCM, "ucm", sd.UCM_GT),
                                             (sd.MEI, "mei", sd.MEI_GT)])
def test_bounds_and_reconstruct_probes(oracle, model, name, intr):
    b = GOLD[f"probe/bounds_{name}"]
    for i in range(len(intr)):
        assert oracle.lib.vgo_lower_bound(model, i) == b[i, 0] and oracle.lib.vgo_upper_bound(model, i) == b[i, 1]
    uv, X, ok = GOLD[f"probe/reconstruct_{name}_uv"], GOLD[f"probe/reconstruct_{name}_X"], GOLD[f"probe/reconstruct_{name}_ok"]
    for i in range(len(uv)):
        Xo, oko = oracle.reconstruct(model, intr, uv[i])
        assert oko == bool(ok[i])
        if oko:
            assert np.abs(Xo - X[i]).max() < 1e-14


@pytest.fixture(scope="module")
def reference():
    from oracle.pyoracle import Reference
    try:
        return Reference()
    except Exception:
        pytest.skip("oracle/_ref is not built (the reference tree only exists in the authoring container)")


@pytest.mark.parametrize("model", [sd.EUCM, sd.UCM, sd.MEI])
def test_oracle_matches_reference_build_live(oracle, reference, model):
    d = sd.make_mono(model, 300, seed=4242 + model)
    args = (model, d["intr_init"], d["board"], d["obs"], [d["xi_init"]], [0], [0])
    a, b = oracle.evaluate_batch(*args, want_H=True, threads=4), reference.evaluate_batch(*args, want_H=True, threads=4)
    assert_close(a["r"], b["r"], "r", PIN_RTOL)
    assert_close(a["J_intr"], b["J_intr"], "J_intr", PIN_RTOL)
    assert_close(a["J_xi"][0], b["J_xi"][0], "J_xi", PIN_RTOL)
    s = sd.make_stereo(100, seed=99)
    args = (sd.EUCM, s["intr2_init"], s["board"], s["obs2"], [s["xi12_init"], s["xi_init"]], [1, 0], [1, 0])
    a, b = oracle.evaluate_batch(*args), reference.evaluate_batch(*args)
    for e in range(2):
        assert_close(a["J_xi"][e], b["J_xi"][e], f"stereo J_xi[{e}]", PIN_RTOL)

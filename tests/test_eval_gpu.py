"""GPU parity of the fused residual + Jacobian + normal-equation kernels against the
CPU oracle (restatement of calib_cost_functions.cpp:28-117), through the C ABI.

Tolerance: 1e-9 of each block's scale (north_star allows 1e-6 relative)."""
import numpy as np
import pytest

import synthdata as sd
from util import RTOL, assert_close, compare_eval, unpack_hessian

pytestmark = pytest.mark.gpu

D, I = 0, 1   # TRANSFORM_DIRECT / TRANSFORM_INVERSE


def make_chain(oracle, xi_gt, status, is_global, seed=99):
    """Random chain whose composition equals xi_gt for every image: all elements are small random
    transforms except the (single) sequence element, which is solved to close the chain."""
    n_img, L = xi_gt.shape[0], len(status)
    seq = [e for e in range(L) if not is_global[e]]
    assert len(seq) == 1
    sq = seq[0]
    u = sd.uniform(seed, L, 6 * L).reshape(L, 6) * 2 - 1
    glob = [u[e] * np.array([0.1, 0.1, 0.1, 0.2, 0.2, 0.2]) for e in range(L)]
    ident = np.zeros(6)

    def chain(elems):
        acc = ident
        for e in elems:
            acc = oracle.compose(acc, glob[e], "compose" if status[e] == D else "compose_inverse")
        return acc
    before, rest = chain(range(sq)), chain(range(sq + 1, L))
    vals = []
    for i in range(n_img):
        y = oracle.compose(before, oracle.compose(xi_gt[i], rest, "compose_inverse"), "inverse_compose")
        vals.append(y if status[sq] == D else oracle.compose(ident, y, "compose_inverse"))
    return [np.ascontiguousarray(np.stack(vals)) if e == sq else glob[e].copy() for e in range(L)]


def both(gpu, oracle, model, intr, board, obs, xis, status, is_global, threads=8):
    g = gpu.eval_chain(model, intr, board, obs, xis, status, is_global, want_H=True)
    o = oracle.evaluate_batch(model, intr, board, obs, xis, status, is_global, want_H=True, threads=threads)
    return g, o


@pytest.mark.parametrize("model", [sd.EUCM, sd.UCM, sd.MEI])
@pytest.mark.parametrize("n_img", [1, 3, 20, 203])
def test_mono_parity(gpu, oracle, model, n_img):
    """C1-shaped problems (and ragged group sizes: 1, 3, 203 are not multiples of the CTA group)."""
    d = sd.make_mono(model, n_img, seed=20241 + model)
    for intr, xi in ((d["intr_gt"], d["xi_gt"]), (d["intr_init"], d["xi_init"])):
        g, o = both(gpu, oracle, model, intr, d["board"], d["obs"], [xi], [D], [0])
        compare_eval(g, o, f"mono model={model} n={n_img}")


@pytest.mark.parametrize("model,n_img", [(sd.EUCM, 10000), (sd.MEI, 10000), (sd.UCM, 10000), (sd.EUCM, 25000)])
def test_full_size_parity(gpu, oracle, model, n_img):
    """C2 / C3 at BASELINE.json's full size, UCM at the same size and C5's per-GPU shard (25 000 images): every
    output element against the oracle."""
    d = sd.make_mono(model, n_img, seed=20242 + model)
    g, o = both(gpu, oracle, model, d["intr_init"], d["board"], d["obs"], [d["xi_init"]], [D], [0])
    compare_eval(g, o, f"full model={model}")
    # size-independent property: the normal-equation block is the Gram matrix of the
    # kernel's own Jacobian/residual rows
    rows = np.concatenate([g["J_intr"], g["J_xi"][0], g["r"][:, :, None]], axis=2)
    gram = np.einsum("nka,nkb->nab", rows, rows)
    Hf = unpack_hessian(g["H"])
    dscale = np.sqrt(np.einsum("nii->ni", gram))
    assert (np.abs(Hf - gram) / (dscale[:, :, None] * dscale[:, None, :])).max() < 1e-12


def test_stereo_full_size_parity(gpu, oracle):
    """C4 at BASELINE.json's full size (5 000 stereo pairs = 540 000 corners): both cameras' blocks against the oracle,
    the two-element chain [xiCam12 INVERSE (global), xiCamBoardStereo DIRECT (sequence)] included."""
    s = sd.make_stereo(5000, seed=20244)
    g, o = both(gpu, oracle, sd.EUCM, s["intr2_init"], s["board"], s["obs2"], [s["xi12_init"], s["xi_init"]], [I, D], [1, 0])
    compare_eval(g, o, "C4 cam2")
    g, o = both(gpu, oracle, sd.EUCM, s["intr1_init"], s["board"], s["obs1"], [s["xi_init"]], [D], [0])
    compare_eval(g, o, "C4 cam1")


def test_stereo_chain_parity(gpu, oracle):
    """C4: camera2 sees the board through [xiCam12 INVERSE (global), xiCamBoard DIRECT (sequence)]
    (data/calib_stereo_example.json:86-92)."""
    s = sd.make_stereo(200, seed=20244)
    for xi12, intr2, xi in ((s["xi12_gt"], s["intr2_gt"], s["xi_gt"]), (s["xi12_init"], s["intr2_init"], s["xi_init"])):
        g, o = both(gpu, oracle, sd.EUCM, intr2, s["board"], s["obs2"], [xi12, xi], [I, D], [1, 0])
        compare_eval(g, o, "stereo cam2")
    g, o = both(gpu, oracle, sd.EUCM, s["intr1_init"], s["board"], s["obs1"], [s["xi_init"]], [D], [0])
    compare_eval(g, o, "stereo cam1")


@pytest.mark.parametrize("model", [sd.EUCM, sd.UCM, sd.MEI])
@pytest.mark.parametrize("status,is_global", [
    ([D, D], [0, 1]), ([I, I], [1, 0]), ([D, I, D], [1, 1, 0]), ([I, D, I, D], [0, 1, 1, 1]),
    ([D, I, D, I, D], [1, 1, 1, 1, 0])])
def test_general_chains(gpu, oracle, model, status, is_global):
    """Every chain length up to the reference's maximum of 5 (unified_calibration.cpp:567), mixed
    DIRECT / INVERSE and global / sequence elements; the chain is built so that the composed
    transform is the ground-truth board pose."""
    n_img = 37
    d = sd.make_mono(model, n_img, seed=777 + len(status))
    xis = make_chain(oracle, d["xi_gt"], status, is_global)
    g, o = both(gpu, oracle, model, d["intr_gt"], d["board"], d["obs"], xis, status, is_global)
    compare_eval(g, o, f"chain {status} {is_global} model={model}")


def test_failed_projection_sentinel(gpu, oracle):
    """eta < 1e-3 and the alpha > 0.5 hemisphere test (eucm.h:46-54): residual 1e15 and zero
    Jacobian rows (calib_cost_functions.cpp:66-70, eucm.h:141-150,198-206)."""
    d = sd.make_mono(sd.EUCM, 16, seed=5)
    xi = d["xi_gt"].copy()
    xi[::2, 2] -= 3.0          # every other board goes behind the camera
    xi[1, 2] = -0.01           # board straddling the validity boundary: mixed corners
    g, o = both(gpu, oracle, sd.EUCM, d["intr_gt"], d["board"], d["obs"], [xi], [D], [0])
    assert (o["r"] == 1e15).any() and (o["r"] != 1e15).any()
    assert ((g["r"] == 1e15) == (o["r"] == 1e15)).all()
    bad = (o["r"] == 1e15)
    assert (g["J_intr"][bad] == 0).all() and (g["J_xi"][0][bad] == 0).all()
    compare_eval(g, o, "sentinel")
    # alpha <= 0.5: only the eta test applies
    intr = d["intr_gt"].copy(); intr[0] = 0.4
    g, o = both(gpu, oracle, sd.EUCM, intr, d["board"], d["obs"], [xi], [D], [0])
    assert ((g["r"] == 1e15) == (o["r"] == 1e15)).all()
    compare_eval(g, o, "sentinel alpha<=0.5")


def test_small_angle_branches(gpu, oracle):
    """theta < 1e-5 (rotationMatrix / interOmegaRot, geometry_core.h:44,162) and the exact zero rotation."""
    d = sd.make_mono(sd.EUCM, 8, seed=11)
    xi = d["xi_gt"].copy()
    xi[:, 3:] = 0.0
    xi[:, 0] = -0.4; xi[:, 1] = -0.25; xi[:, 2] = 0.8
    xi[1, 3:] = [3e-6, -2e-6, 1e-6]
    xi[2, 3:] = [9.9e-6, 0, 0]
    xi[3, 3:] = [1.01e-5, 0, 0]
    xi[4, 3:] = [1e-7, 1e-7, -1e-7]
    xi[5, 3:] = [2e-3, -1e-3, 5e-4]
    g, o = both(gpu, oracle, sd.EUCM, d["intr_gt"], d["board"], d["obs"], [xi], [D], [0])
    # the reference routes the pose through a quaternion round trip whose own small-angle
    # switch (1e-6 / 1e-5, quaternion.h:34,88) perturbs R by O(theta^2) ~ 1e-10: allow 1e-8 here
    compare_eval(g, o, "small angles", rtol=1e-8)


@pytest.mark.parametrize("nx,ny", [(2, 2), (8, 5), (12, 9), (20, 15), (36, 28)])
def test_board_sizes(gpu, oracle, nx, ny):
    """4-point IR grids (unified_calibration.cpp:234-250), the 8x5 board of data/calib_example.json,
    and boards large enough to need a single image per CTA and a looped corner phase."""
    d = sd.make_mono(sd.EUCM, 9, seed=nx * 100 + ny, nx=nx, ny=ny, size=0.8 / max(nx - 1, 1) * 0.9)
    g, o = both(gpu, oracle, sd.EUCM, d["intr_init"], d["board"], d["obs"], [d["xi_init"]], [D], [0])
    compare_eval(g, o, f"board {nx}x{ny}")


def test_optional_outputs_and_empty(gpu, oracle):
    d = sd.make_mono(sd.EUCM, 5, seed=3)
    args = (sd.EUCM, d["intr_init"], d["board"], d["obs"], [d["xi_init"]], [D], [0])
    full = gpu.eval_chain(*args, want_H=True)
    only_r = gpu.eval_chain(*args, want_J=False)
    assert only_r["J_intr"] is None and (only_r["r"] == full["r"]).all()
    only_h = gpu.eval_chain(*args, want_r=False, want_J=False, want_H=True)
    assert (only_h["H"] == full["H"]).all()
    empty = gpu.eval_chain(sd.EUCM, d["intr_init"], d["board"], np.zeros((0, 2 * d["P"])),
                           [np.zeros((0, 6))], [D], [0], want_H=True)
    assert empty["r"].shape == (0, 2 * d["P"]) and empty["H"].shape[0] == 0


def test_argument_errors(gpu):
    d = sd.make_mono(sd.EUCM, 2, seed=3)
    big = sd.make_mono(sd.EUCM, 1, seed=4, nx=40, ny=30, size=0.02)
    with pytest.raises(gpu.VisgeomError, match="too many points"):   # 1200 corners exceed one CTA's 227 KB
        gpu.eval_chain(sd.EUCM, big["intr_init"], big["board"], big["obs"], [big["xi_init"]], [D], [0])
    with pytest.raises(gpu.VisgeomError):   # chain longer than 5, unified_calibration.cpp:567
        gpu.eval_chain(sd.EUCM, d["intr_init"], d["board"], d["obs"], [d["xi_init"]] * 6, [D] * 6, [0] * 6)
    with pytest.raises(gpu.VisgeomError):   # invalid camera model name, :177
        gpu.eval_chain(7, d["intr_init"], d["board"], d["obs"], [d["xi_init"]], [D], [0])


def test_deterministic(gpu):
    d = sd.make_mono(sd.EUCM, 257, seed=9)
    args = (sd.EUCM, d["intr_init"], d["board"], d["obs"], [d["xi_init"]], [D], [0])
    a = gpu.eval_chain(*args, want_H=True)
    b = gpu.eval_chain(*args, want_H=True)
    for k in ("r", "J_intr", "H"):
        assert (a[k] == b[k]).all()


def test_concurrent_host_calls(gpu, oracle):
    """vg_eval_chain from four host threads at once (Ceres evaluates residual blocks from num_threads threads; every
    calling thread owns its streams, pinned slots and device workspace): each thread's outputs against the oracle, on
    problems of different models and sizes, repeated so that the calls really overlap."""
    import threading
    jobs = []
    for t, (model, n_img) in enumerate([(sd.EUCM, 700), (sd.MEI, 450), (sd.UCM, 900), (sd.EUCM, 33)]):
        d = sd.make_mono(model, n_img, seed=600 + t)
        want = oracle.evaluate_batch(model, d["intr_init"], d["board"], d["obs"], [d["xi_init"]], [D], [0], want_H=True)
        jobs.append((model, d, want))
    errors = []

    def work(model, d, want):
        try:
            for _ in range(6):
                g = gpu.eval_chain(model, d["intr_init"], d["board"], d["obs"], [d["xi_init"]], [D], [0], want_H=True)
                compare_eval(g, want, f"thread model={model}")
        except Exception as e:       # noqa: BLE001 -- reported below, in the main thread
            errors.append(repr(e))
    threads = [threading.Thread(target=work, args=j) for j in jobs]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors


def test_host_call_chunking(gpu, oracle, monkeypatch):
    """The chunked host path (chunk boundaries, the ragged last chunk, helper copy threads, a stereo chain whose global
    element lives in the fixed inputs): 3 000 images cut into many small chunks, every element against the oracle."""
    s = sd.make_stereo(3000, seed=611)
    args = (sd.EUCM, s["intr2_init"], s["board"], s["obs2"], [s["xi12_init"], s["xi_init"]], [I, D], [1, 0])
    g = gpu.eval_chain(*args, want_H=True)
    o = oracle.evaluate_batch(*args, want_H=True, threads=8)
    compare_eval(g, o, "chunked stereo")
    g2 = gpu.eval_chain(*args, want_H=True, want_J=False)
    assert (g2["r"] == g["r"]).all() and (g2["H"] == g["H"]).all()


def test_registered_output_arrays_are_written_directly(gpu):
    """vg_host_register: page-locked output arrays take the DMA directly (no staging slot, no host memcpy); every output
    independently -- all registered, only some, none -- and the bytes are the same in every case."""
    s = sd.make_stereo(3000, seed=612)
    args = (sd.EUCM, s["intr2_init"], s["board"], s["obs2"], [s["xi12_init"], s["xi_init"]], [I, D], [1, 0])
    plain = gpu.eval_chain(*args, want_H=True)
    out = gpu.eval_chain(*args, want_H=True)
    for a in (out["r"], out["J_intr"], out["H"], *out["J_xi"]):
        a[...] = -1.0
    regs = [out["r"], out["J_intr"], out["H"], *out["J_xi"]]
    for a in regs:
        gpu.host_register(a)
    try:
        gpu.eval_chain(*args, want_H=True, out=out)
        for k in ("r", "J_intr", "H"):
            assert np.array_equal(out[k], plain[k]), k
        for e in range(2):
            assert np.array_equal(out["J_xi"][e], plain["J_xi"][e]), e
    finally:
        for a in regs:
            gpu.host_unregister(a)
    # only the largest array registered: the others still go through the staging slots
    out["J_intr"][...] = -1.0; out["r"][...] = -1.0
    gpu.host_register(out["J_intr"])
    try:
        gpu.eval_chain(*args, want_H=True, out=out)
        assert np.array_equal(out["J_intr"], plain["J_intr"]) and np.array_equal(out["r"], plain["r"])
        assert np.array_equal(out["J_xi"][1], plain["J_xi"][1])
    finally:
        gpu.host_unregister(out["J_intr"])
    with pytest.raises(gpu.VisgeomError):
        gpu.host_unregister(out["r"])                      # not registered: the CUDA error comes back as a message

"""Checkerboard detector after the response stage (corner_detector.cpp:223-260, :331-1298), CPU side:
  * the checker itself: the reference's own corner_detector.cpp compiled where it lies (oracle/_ref, only where
    /root/reference exists) reproduces the committed fixtures (tests/golden/detector.npz), its -O2 and -O0 builds agree,
    the C oracle's response maps equal the reference build's, and the restated line-search minimiser of the Ceres
    stand-in ends where an independent minimiser (scipy BFGS) ends;
  * the product's host stages (vg_detector_host_stages: candidate tests, flood-fill graph, pattern search, initPoin,
    improveCorners' reach -- no GPU involved) against the fixtures: candidates in graph order, the grid, the start
    values, all exactly."""
import os

import numpy as np
import pytest

from util import local_maxima_np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "detector.npz"))
IMAGES = GOLD["images"]
SIGMAS = tuple(GOLD["sigmas"])
N = len(IMAGES)
HAVE_REF = os.path.isdir("/root/reference/include") or os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libvisgeom_refdet.so"))
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built (needs /root/reference)")


@pytest.fixture(scope="module")
def refdet():
    from oracle.pyoracle import ReferenceDetector
    return ReferenceDetector()


def test_fixture_covers_the_cases():
    found = [bool(GOLD[f"{k}/found"]) for k in range(N)]
    assert sum(found) >= 6 and found.count(False) >= 3                      # boards, a cut board, noise, a flat image
    late = [k for k in range(N) if found[k] and len(GOLD[f"{k}/scale0/pattern"]) != 54]
    assert len(late) >= 2                                                   # the later scales of detectPattern are exercised
    for k in range(N):
        if found[k]:                                                        # the reference finds the rendered corners
            assert np.abs(GOLD[f"{k}/refined"] - GOLD[f"{k}/true"]).max() < 0.35
            assert np.abs(GOLD[f"{k}/grid"] - GOLD[f"{k}/true"]).max() < 1.0


@needs_ref
@pytest.mark.parametrize("k", range(N))
def test_reference_build_reproduces_the_fixture(refdet, k):
    ok, refined, start, iters = refdet.detect_pattern(IMAGES[k], improve=True)
    assert ok == bool(GOLD[f"{k}/found"])
    if ok:
        assert np.array_equal(refined, GOLD[f"{k}/refined"]) and np.array_equal(start, GOLD[f"{k}/start"])
        assert np.array_equal(iters, GOLD[f"{k}/iters"])
    for s, sigma in enumerate(SIGMAS):
        st = refdet.stages(IMAGES[k], sigma)
        assert np.array_equal(st["cand"], GOLD[f"{k}/scale{s}/cand"]) and np.array_equal(st["pattern"], GOLD[f"{k}/scale{s}/pattern"])


@needs_ref
def test_reference_O0_build_equals_O2_build_with_the_loop_written_out(refdet):
    from oracle.pyoracle import ReferenceDetector
    exact = ReferenceDetector(exact=True)          # detectPattern as written, improveCorners through initPoin
    for k in (0, 5):
        a = refdet.detect_pattern(IMAGES[k], improve=True)
        b = exact.detect_pattern(IMAGES[k], improve=True)
        assert a[0] and b[0] and np.array_equal(a[1], b[1])


@needs_ref
@pytest.mark.parametrize("k", (0, 4, 8))
def test_oracle_response_equals_reference_build(oracle, refdet, k):
    for sigma in SIGMAS:
        o, m = oracle.corner_response(IMAGES[k], 0.7, sigma), refdet.maps(IMAGES[k], sigma)
        for key in ("resp", "gradx", "grady", "imgrad"):
            assert np.array_equal(o[key], m[key]), (key, sigma)
        n2 = 1 + 2 * int(np.ceil(sigma))
        assert np.array_equal(oracle.gaussian_blur_u8(IMAGES[k], 3, 0.7), m["s1"])
        assert np.array_equal(oracle.gaussian_blur_u8(IMAGES[k], n2, sigma), m["s2"])


@needs_ref
def test_restated_minimiser_ends_where_scipy_ends(oracle, refdet):
    """The Ceres stand-in's L-BFGS + Wolfe search (oracle/shim/ceres/gradient_solver.h) cannot be pinned on Ceres; hold it
    against an independent minimiser of the reference's own SubpixelCorner cost instead."""
    from scipy.optimize import minimize
    m = oracle.corner_response(IMAGES[0], 0.7, 1.4)
    worst = 0.0
    for i in (0, 13, 27, 40, 53):
        prior = GOLD["0/grid"][i].astype(np.float64)
        x0 = GOLD["0/start"][i]
        xs, it, cost = refdet.subpixel_solve(m["gradx"], m["grady"], prior, 5.0, x0)
        res = minimize(lambda x: refdet.subpixel_evaluate(m["gradx"], m["grady"], prior, 5.0, x), x0, jac=True, method="BFGS",
                       options=dict(gtol=1e-9))
        assert cost <= refdet.subpixel_evaluate(m["gradx"], m["grady"], prior, 5.0, x0)[0]
        worst = max(worst, np.abs(xs[:2] - res.x[:2]).max())
        assert abs(cost - res.fun) <= 2e-6 * abs(res.fun) + 1e-9            # stops on Ceres' function tolerance 1e-6
    assert worst < 2e-2, worst


def _gpu_stage_stand_in(oracle, img, sigma):
    """What the GPU hands the host stages, from the CPU oracle: the two blurred images and the local maxima."""
    r = oracle.corner_response(img, 0.7, sigma)
    R = int(round(1.5 * sigma))
    val, uv = local_maxima_np(r["resp"], float(r["avg"]), R)
    s1 = oracle.gaussian_blur_u8(img, 3, 0.7)
    s2 = oracle.gaussian_blur_u8(img, 1 + 2 * int(np.ceil(sigma)), sigma)
    return s1, s2, val, uv, R


@pytest.mark.parametrize("k", range(N))
def test_host_stages_equal_the_reference(vg, oracle, k):
    img = IMAGES[k]
    hit = False
    for s, sigma in enumerate(SIGMAS):
        s1, s2, val, uv, R = _gpu_stage_stand_in(oracle, img, sigma)
        perm = np.random.default_rng(k).permutation(len(val))               # the GPU emits the maxima in any order
        r = vg.detector_host_stages(img, s1, s2, val[perm], uv[perm], R)
        cand = GOLD[f"{k}/scale{s}/cand"]
        assert r["cand"].shape == cand.shape and np.array_equal(r["cand"], cand), (k, sigma)
        pat = GOLD[f"{k}/scale{s}/pattern"]
        assert r["found"] == (len(pat) == 54)
        if r["found"]:
            assert np.array_equal(r["grid"], cand[pat])
        if r["found"] and not hit:                                          # the scale detectPattern stops at
            hit = True
            assert np.array_equal(r["grid"], GOLD[f"{k}/grid"])
            assert np.array_equal(r["start"], GOLD[f"{k}/start"])           # initPoin, bit for bit
    assert hit == bool(GOLD[f"{k}/found"])


def test_host_stages_reject_bad_arguments(vg):
    img = IMAGES[0]
    with pytest.raises(vg.VisgeomError):
        vg.detector_host_stages(img[:4, :4], img[:4, :4], img[:4, :4], [], np.zeros((0, 2)), 2)
    with pytest.raises(vg.VisgeomError):
        vg.detector_host_stages(img, img, img, [], np.zeros((0, 2)), 0)
    r = vg.detector_host_stages(img, img, img, [], np.zeros((0, 2)), 2)     # no maxima: no pattern, no error
    assert not r["found"] and len(r["cand"]) == 0


def test_detect_pattern_fails_loudly_without_a_gpu(vg):
    if vg.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(vg.VisgeomError):
        vg.detect_pattern(IMAGES[0])
    with pytest.raises(vg.VisgeomError):
        vg.subpixel_refine(np.zeros((8, 8), np.float32), np.zeros((8, 8), np.float32), [[4, 4]], [3.0], [[4, 4, 0, 1.5, 0]])


# ---- the Python restatement (oracle/detector_oracle.py): pinned on the fixtures, then used as the checker on new pictures ----
def _python_oracle(oracle, img, sigma):
    from oracle.detector_oracle import DetectorOracle
    m = oracle.corner_response(img, 0.7, float(sigma))
    s2 = oracle.gaussian_blur_u8(img, 1 + 2 * int(np.ceil(sigma)), float(sigma))
    return DetectorOracle(img, m, s2, 9, 6, int(round(1.5 * sigma)))


@pytest.mark.parametrize("k", range(N))
def test_python_restatement_reproduces_the_fixture(oracle, k):
    """candidates in graph order (the standard library's heap order among equal keys included), the pattern, initPoin's
    values: exactly what the reference build recorded"""
    first = True
    for s, sigma in enumerate(SIGMAS):
        D = _python_oracle(oracle, IMAGES[k], sigma)
        cand, pat = D.detect()
        assert np.array_equal(np.array(cand, dtype=np.int32).reshape(-1, 2), GOLD[f"{k}/scale{s}/cand"]), (k, s)
        assert list(pat) == list(GOLD[f"{k}/scale{s}/pattern"]), (k, s)
        if len(pat) == 54 and first:
            first = False
            grid = [cand[i] for i in pat]
            assert np.array_equal(np.array(grid), GOLD[f"{k}/grid"])
            assert np.array_equal(np.array([D.init_point(p) for p in grid]), GOLD[f"{k}/start"])


def test_python_restatement_of_the_subpixel_cost(oracle):
    from oracle.detector_oracle import subpixel_evaluate
    m = oracle.corner_response(IMAGES[0], 0.7, 1.4)
    for pr, x, ln, c, g in zip(GOLD["eval/prior"], GOLD["eval/x"], GOLD["eval/length"], GOLD["eval/cost"], GOLD["eval/grad"]):
        cc, gg = subpixel_evaluate(m["gradx"], m["grady"], pr, float(ln), x)
        assert abs(cc - c) <= 1e-12 * max(1.0, abs(c)) and np.abs(gg - g).max() <= 1e-12 * max(1.0, np.abs(g).max())


def test_host_stages_equal_the_python_restatement_on_new_pictures(vg, oracle):
    """pictures that are in no fixture (other sizes, seeds, noise levels; a board cut by the border), checked without the
    reference build: the product's host stages against the Python restatement"""
    import synthdata as sd
    from oracle.detector_oracle import refinement_reach
    rng = np.random.default_rng(99)
    found = 0
    for trial in range(8):
        w = int(rng.integers(180, 420)); h = int(w * rng.uniform(0.65, 0.85))
        img, _ = sd.render_board_image(w, h, seed=61000 + trial, model=(sd.EUCM, sd.MEI, sd.UCM)[trial % 3],
                                       noise=float(rng.choice([1.0, 5.0, 11.0])), supersample=2)
        if trial == 5:
            img = np.ascontiguousarray(img[:, : w * 3 // 5])
        for sigma in SIGMAS:
            D = _python_oracle(oracle, img, sigma)
            cand, pat = D.detect()
            s1, s2, val, uv, R = _gpu_stage_stand_in(oracle, img, sigma)
            r = vg.detector_host_stages(img, s1, s2, val, uv, R)
            assert np.array_equal(r["cand"], np.array(cand, dtype=np.int32).reshape(-1, 2)), (trial, sigma)
            assert r["found"] == (len(pat) == 54), (trial, sigma)
            if r["found"]:
                found += 1
                grid = np.array([cand[i] for i in pat])
                assert np.array_equal(r["grid"], grid)
                assert np.array_equal(r["start"], np.array([D.init_point(tuple(p)) for p in grid]))
                assert np.array_equal(r["reach"], refinement_reach(grid, 9))
    assert found >= 10


@pytest.mark.parametrize("k,sigma", [(0, 1.4), (5, 1.0)])
def test_python_restatement_of_the_refinement(oracle, k, sigma):
    """improveCorners through the Python restatement of the minimiser: the reference build's refined corners and
    iteration counts, bit for bit (the two restatements of Ceres' line search walk the same path)"""
    from oracle.detector_oracle import improve_corners
    m = oracle.corner_response(IMAGES[k], 0.7, sigma)
    refined, its = improve_corners(m["gradx"], m["grady"], GOLD[f"{k}/grid"], GOLD[f"{k}/start"], 9)
    assert np.array_equal(refined, GOLD[f"{k}/refined"]) and np.array_equal(its, GOLD[f"{k}/iters"])


@needs_ref
def test_python_restatement_against_the_reference_build_on_random_pictures(oracle, refdet):
    """beyond the fixtures: a dozen random pictures (three models, noise up to 12 grey levels, a cut board, pure noise),
    every scale: candidates in graph order and the pattern, the Python restatement against the reference build"""
    import synthdata as sd
    rng = np.random.default_rng(2024)
    found = 0
    for trial in range(12):
        w = int(rng.integers(160, 420)); h = int(w * rng.uniform(0.6, 0.9))
        img, _ = sd.render_board_image(w, h, seed=63000 + trial, model=(sd.EUCM, sd.MEI, sd.UCM)[trial % 3],
                                       noise=float(rng.choice([0.5, 4.0, 12.0])), supersample=2)
        if trial == 4:
            img = np.ascontiguousarray(img[:, : w // 2])
        if trial == 9:
            img = rng.integers(0, 256, img.shape, dtype=np.uint8)
        for sigma in SIGMAS:
            st = refdet.stages(img, sigma)
            cand, pat = _python_oracle(oracle, img, sigma).detect()
            assert np.array_equal(np.array(cand, dtype=np.int32).reshape(-1, 2), st["cand"]), (trial, sigma)
            assert list(pat) == list(st["pattern"]), (trial, sigma)
            found += len(pat) == 54
    assert found >= 15

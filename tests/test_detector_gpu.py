"""Checkerboard detector on the GPU (vg_detect_pattern, vg_subpixel_evaluate, vg_subpixel_refine) against the fixtures
recorded from the reference's own corner_detector.cpp (tests/golden/detector.npz):
  * found / not found and the integer grid: exactly;
  * SubpixelCorner::Evaluate (cost and gradient): 1e-11 relative (the 28 samples are summed in another order and
    sin / cos come from another library);
  * refined corners: 1e-4 pixel.  The minimiser stops on Ceres' relative function tolerance of 1e-6, where last-bit
    differences of the cost can move the final iterate by ~1e-6 pixel; the reference's own result is ~1e-2 pixel from
    the rendered truth, so the tolerance is 100x below what the detector resolves.  (The minimiser itself is a
    restatement of Ceres' -- tests/test_detector_oracle.py holds it against scipy.)
  * at full size (1280 x 800, a batch of boards): every board found, refined corners within 0.3 pixel of the rendered
    truth, and -- where oracle/_ref travelled to the box -- equal to the reference build's."""
import os

import numpy as np
import pytest

import synthdata as sd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "detector.npz"))
IMAGES = GOLD["images"]
N = len(IMAGES)
REFINE_TOL = 1e-4
pytestmark = pytest.mark.gpu


def test_grid_without_refinement_is_exact(gpu):
    found, corners = gpu.detect_pattern(IMAGES, improve=False)
    for k in range(N):
        assert found[k] == bool(GOLD[f"{k}/found"]), k
        if found[k]:
            assert np.array_equal(corners[k], GOLD[f"{k}/grid"].astype(np.float64)), k
        else:
            assert np.isnan(corners[k]).all()


def test_refined_corners_match_the_reference(gpu):
    found, corners = gpu.detect_pattern(IMAGES, improve=True)
    worst = 0.0
    for k in range(N):
        assert found[k] == bool(GOLD[f"{k}/found"]), k
        if found[k]:
            worst = max(worst, np.abs(corners[k] - GOLD[f"{k}/refined"]).max())
    assert worst < REFINE_TOL, worst


def test_single_image_and_batch_agree(gpu):
    fb, cb = gpu.detect_pattern(IMAGES[:3], improve=True)
    for k in range(3):
        f1, c1 = gpu.detect_pattern(IMAGES[k], improve=True)
        assert f1 == fb[k] and np.array_equal(c1, cb[k])


def test_subpixel_cost_and_gradient(gpu, oracle):
    m = oracle.corner_response(IMAGES[0], 0.7, 1.4)
    cost, grad = gpu.subpixel_evaluate(m["gradx"], m["grady"], GOLD["eval/prior"], GOLD["eval/length"], GOLD["eval/x"])
    assert np.abs(cost - GOLD["eval/cost"]).max() <= 1e-11 * np.abs(GOLD["eval/cost"]).max()
    assert np.abs(grad - GOLD["eval/grad"]).max() <= 1e-11 * np.abs(GOLD["eval/grad"]).max()


def test_refinement_alone_from_the_reference_start_values(gpu, oracle):
    m = oracle.corner_response(IMAGES[0], 0.7, 1.4)
    grid = GOLD["0/grid"].astype(np.float64)
    reach = np.empty(54)
    for i in range(54):                                     # improveCorners' radMax (corner_detector.cpp:164-175)
        a = grid[i - 9] if i > 9 else grid[i + 9]
        b = grid[i - 1] if i > 0 else grid[i + 1]
        reach[i] = min(7.0, np.linalg.norm(grid[i] - a) * 0.7, np.linalg.norm(grid[i] - b) * 0.7)
    refined, iters = gpu.subpixel_refine(m["gradx"], m["grady"], grid, reach, GOLD["0/start"])
    assert np.abs(refined - GOLD["0/refined"]).max() < REFINE_TOL
    assert np.abs(iters - GOLD["0/iters"]).max() <= 1 and (iters == GOLD["0/iters"]).mean() > 0.9


def test_empty_batch_and_bad_arguments(gpu):
    found, corners = gpu.detect_pattern(np.zeros((0, 240, 320), np.uint8))
    assert found.shape == (0,) and corners.shape == (0, 54, 2)
    with pytest.raises(gpu.VisgeomError):
        gpu.detect_pattern(np.zeros((1, 4, 4), np.uint8))
    with pytest.raises(gpu.VisgeomError):
        gpu.detect_pattern(IMAGES[:1], nx=1, ny=6)


def test_other_board_sizes(gpu):
    img, uv = sd.render_board_image(400, 300, seed=20260, nx=7, ny=5, model=sd.EUCM)
    found, c = gpu.detect_pattern(img, nx=7, ny=5)
    assert found and np.abs(c - uv).max() < 0.3
    found, _ = gpu.detect_pattern(img, nx=9, ny=6)          # the wrong board is not found
    assert not found


def test_full_size_batch(gpu, monkeypatch):
    """BASELINE-size images (1280 x 800), all three camera models, in passes of 5 images (VG_DETECT_CHUNK)."""
    monkeypatch.setenv("VG_DETECT_CHUNK", "5")
    imgs, truth = [], []
    for k in range(12):
        img, uv = sd.render_board_image(1280, 800, seed=20300 + k, model=(sd.EUCM, sd.MEI, sd.UCM)[k % 3], supersample=2)
        imgs.append(img); truth.append(uv)
    imgs = np.stack(imgs)
    found, corners = gpu.detect_pattern(imgs, improve=True)
    assert found.sum() >= 10                                # the reference misses one of these boards; so must we
    f0, g0 = gpu.detect_pattern(imgs, improve=False)
    assert np.array_equal(f0, found)
    for k in np.nonzero(found)[0]:
        assert np.abs(g0[k] - truth[k]).max() < 2.5, k      # the integer grid is the rendered grid (the reference's own
                                                            # maxima sit up to ~1.6 px off on the most oblique boards)
        assert np.abs(corners[k] - g0[k]).max() < 2.0, k    # and the refinement stays near it
    assert np.median([np.abs(corners[k] - truth[k]).max() for k in np.nonzero(found)[0]]) < 0.3
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libvisgeom_refdet.so")):
        from oracle.pyoracle import ReferenceDetector
        ref = ReferenceDetector()
        for k in range(12):
            ok, grid, _, _ = ref.detect_pattern(imgs[k], improve=False)
            assert ok == found[k], k
            if ok:
                assert np.array_equal(g0[k], grid), k
                ok, refined, _, _ = ref.detect_pattern(imgs[k], improve=True)
                assert np.abs(corners[k] - refined).max() < REFINE_TOL, k


def test_random_pictures_against_the_reference_build(gpu):
    """Thirty pictures of random size, camera model and noise level (boards cut by the border and pure noise among them),
    each held against the reference build run on the box: found / not found, the integer grid exactly, the refined
    corners to the tolerance above.  Needs oracle/_ref (it travels with the snapshot)."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libvisgeom_refdet.so")):
        pytest.skip("oracle/_ref not present on this box")
    from oracle.pyoracle import ReferenceDetector
    ref = ReferenceDetector()
    rng = np.random.default_rng(77)
    n_found = 0
    for trial in range(30):
        w = int(rng.integers(200, 1000)); h = int(w * rng.uniform(0.6, 0.9))
        img, _ = sd.render_board_image(w, h, seed=50000 + trial, model=(sd.EUCM, sd.MEI, sd.UCM)[trial % 3],
                                       noise=float(rng.choice([0.5, 2.0, 6.0, 12.0])), supersample=2)
        if trial % 7 == 3:
            img = np.ascontiguousarray(img[:, : w * 2 // 3])
        if trial % 11 == 5:
            img = rng.integers(0, 256, img.shape, dtype=np.uint8)
        ok, grid, _, _ = ref.detect_pattern(img, improve=False)
        f0, g0 = gpu.detect_pattern(img, improve=False)
        assert f0 == ok, trial
        if ok:
            n_found += 1
            assert np.array_equal(g0, grid), trial
            _, refined, _, _ = ref.detect_pattern(img, improve=True)
            f1, c1 = gpu.detect_pattern(img, improve=True)
            assert f1 and np.abs(c1 - refined).max() < REFINE_TOL, (trial, np.abs(c1 - refined).max())
    assert n_found >= 15


def test_random_pictures_against_the_python_restatement(gpu, oracle):
    """The same kind of check with the checker that always travels (oracle/detector_oracle.py, pinned on the reference
    build's fixtures): the whole GPU + host pipeline's found flag and integer grid on pictures of no fixture, scale by
    scale as detectPattern tries them."""
    from oracle.detector_oracle import DetectorOracle, improve_corners
    rng = np.random.default_rng(123)
    n_found = 0
    for trial in range(8):
        w = int(rng.integers(200, 480)); h = int(w * rng.uniform(0.65, 0.85))
        img, _ = sd.render_board_image(w, h, seed=62000 + trial, model=(sd.EUCM, sd.MEI, sd.UCM)[trial % 3],
                                       noise=float(rng.choice([1.0, 6.0, 13.0])), supersample=2)
        if trial == 3:
            img = np.ascontiguousarray(img[:, : w * 3 // 5])
        want = None
        for sigma in (1.4, 2.0, 1.0):                       # corner_detector.cpp:225-245
            m = oracle.corner_response(img, 0.7, sigma)
            s2 = oracle.gaussian_blur_u8(img, 1 + 2 * int(np.ceil(sigma)), sigma)
            D = DetectorOracle(img, m, s2, 9, 6, int(round(1.5 * sigma)))
            cand, pat = D.detect()
            if len(pat) == 54:
                want = np.array([cand[i] for i in pat], dtype=np.float64)
                break
        found, grid = gpu.detect_pattern(img, improve=False)
        assert found == (want is not None), trial
        if found:
            n_found += 1
            assert np.array_equal(grid, want), trial
            if n_found <= 3:                                # and the refinement, through the restated minimiser
                start = np.array([D.init_point((int(p[0]), int(p[1]))) for p in want])
                refined, _ = improve_corners(m["gradx"], m["grady"], want, start, 9)
                _, got = gpu.detect_pattern(img, improve=True)
                assert np.abs(got - refined).max() < REFINE_TOL, (trial, np.abs(got - refined).max())
    assert n_found >= 4

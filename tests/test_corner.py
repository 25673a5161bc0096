"""First stage of the checkerboard detector (CornerDetector::computeResponse, corner_detector.cpp:262-329):
  * CPU: the oracle's blur against OpenCV's own output (fixtures recorded from cv2 by tests/golden/make_corner_golden.py),
    bit for bit, and the response map's basic properties on rendered boards;
  * GPU: vg_corner_response against the oracle -- every float of the four maps bit for bit, the mean to 1e-12."""
import os

import numpy as np
import pytest

import synthdata as sd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "corner_response.npz"))
NAMES = sorted({k.split("/")[0] for k in GOLD.files if "/" in k})
SIGMAS = (1.4, 2.0, 1.0)            # detectPattern's schedule (corner_detector.cpp:229)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_blur_is_opencv_bit_for_bit(oracle, name):
    img = GOLD[f"{name}/img"]
    for key in [k for k in GOLD.files if k.startswith(f"{name}/blur_")]:
        n, s = key.split("blur_")[1].split("_")
        assert (oracle.gaussian_blur_u8(img, int(n), float(s)) == GOLD[key]).all(), key


def test_oracle_response_peaks_at_the_board_corners(oracle):
    img, uv = GOLD["eucm_173x131/img"], GOLD["eucm_173x131/corners"]
    r = oracle.corner_response(img, 0.7, 1.4)
    assert r["count"] == int((r["resp"] > 0).sum()) and r["avg"] > 0
    assert (r["resp"][0] == 0).all() and (r["resp"][:, -1] == 0).all()          # one-pixel border stays zero (:272-273)
    ys, xs = np.unravel_index(np.argsort(r["resp"].ravel())[-len(uv):], r["resp"].shape)
    d = np.min(np.hypot(xs[:, None] - uv[None, :, 0], ys[:, None] - uv[None, :, 1]), axis=1)
    assert d.max() < 1.5                                                        # the strongest saddles ARE the corners


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_response_matches_oracle_bit_for_bit(gpu, oracle, name):
    img = GOLD[f"{name}/img"]
    for s2 in SIGMAS:
        g = gpu.corner_response(img, 0.7, s2)
        o = oracle.corner_response(img, 0.7, s2)
        for k in ("resp", "gradx", "grady", "imgrad"):
            assert (g[k].view(np.uint32) == o[k].view(np.uint32)).all(), (name, s2, k)
        assert g["count"][0] == o["count"]
        assert abs(g["avg"][0] - o["avg"]) <= 1e-12 * o["avg"]


@pytest.mark.gpu
def test_gpu_response_batch_full_size(gpu, oracle):
    """A batch of 1280 x 800 images (the calibration image size, data/calib_example.json) in one launch; two of them
    against the oracle, the batch against itself image by image."""
    imgs = np.stack([sd.render_board_image(1280, 800, seed=300 + i, supersample=1)[0] for i in range(3)])
    g = gpu.corner_response(imgs, 0.7, 1.4)
    for i in (0, 2):
        o = oracle.corner_response(imgs[i], 0.7, 1.4)
        for k in ("resp", "gradx", "grady", "imgrad"):
            assert (g[k][i].view(np.uint32) == o[k].view(np.uint32)).all(), (i, k)
        assert g["count"][i] == o["count"], (g["count"][i], o["count"])
        assert abs(g["avg"][i] - o["avg"]) <= 1e-11 * o["avg"], (g["avg"][i], o["avg"])
    one = gpu.corner_response(imgs[1], 0.7, 1.4)
    assert (one["resp"] == g["resp"][1]).all() and one["avg"][0] == g["avg"][1]


@pytest.mark.gpu
def test_gpu_response_argument_errors(gpu):
    img = np.zeros((8, 8), dtype=np.uint8)
    with pytest.raises(gpu.VisgeomError, match="sigma out of range"):
        gpu.corner_response(img, 0.7, 5.0)
    with pytest.raises(gpu.VisgeomError, match="bad image size"):
        gpu.corner_response(np.zeros((2, 8), dtype=np.uint8), 0.7, 1.4)
    r = gpu.corner_response(img, 0.7, 1.4)          # a flat image: nothing kept, the mean is 0 / 0 as in the reference
    assert r["count"][0] == 0 and np.isnan(r["avg"][0]) and (r["resp"] == 0).all()

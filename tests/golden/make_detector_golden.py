"""Golden vectors for the checkerboard detector, from the REFERENCE's own corner_detector.cpp compiled where it lies
(oracle/_ref/libvisgeom_refdet.so: `make -C oracle ref`; OpenCV / Ceres are stand-ins, see oracle/shim):
    python tests/golden/make_detector_golden.py
Renders synthetic board images (synthdata.render_board_image: three camera models, several noise levels, a board cut by
the image border, pure noise, images whose first scale fails) and stores per image: found, the integer grid, the refined
corners, initPoin's start values, the minimiser's iteration counts, and per scale the candidates in graph order and the
mean response; plus SubpixelCorner::Evaluate at perturbed parameter vectors -> tests/golden/detector.npz."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import synthdata as sd  # noqa: E402
from oracle.pyoracle import ReferenceDetector  # noqa: E402

W, H = 320, 240
CASES = [  # (seed, model, noise, variant)
    (20250, sd.EUCM, 2.0, ""), (20251, sd.MEI, 2.0, ""), (20253, sd.UCM, 0.5, ""), (30007, sd.EUCM, 6.0, ""),
    (40016, sd.UCM, 14.0, ""), (40084, sd.MEI, 8.0, ""), (40100, sd.UCM, 14.0, ""),      # found at the third scale only
    (30011, sd.UCM, 2.0, "cut"), (30012, sd.EUCM, 2.0, "noise"), (30013, sd.EUCM, 2.0, "flat"),
]
SIGMAS = (1.4, 2.0, 1.0)


def render(seed, model, noise, variant):
    img, uv = sd.render_board_image(W, H, seed=seed, model=model, noise=noise, supersample=2)
    if variant == "cut":                       # everything right of the board's centre replaced by background
        img = img.copy(); img[:, int(uv[:, 0].mean()):] = 120
    elif variant == "noise":
        img = np.random.default_rng(seed).integers(0, 256, img.shape, dtype=np.uint8)
    elif variant == "flat":
        img = np.full_like(img, 97)
    return np.ascontiguousarray(img), uv


def main():
    ref = ReferenceDetector()
    exact = ReferenceDetector(exact=True)
    blob = {"size": np.array([W, H]), "sigmas": np.array(SIGMAS)}
    imgs = []
    rng = np.random.default_rng(11)
    for k, (seed, model, noise, variant) in enumerate(CASES):
        img, uv = render(seed, model, noise, variant)
        imgs.append(img)
        ok0, grid, _, _ = ref.detect_pattern(img, improve=False)
        ok, refined, start, iters = ref.detect_pattern(img, improve=True)
        ok_x, refined_x, _, _ = exact.detect_pattern(img, improve=True)
        assert ok == ok0 == ok_x and (not ok or np.array_equal(refined, refined_x)), "the -O2 and -O0 builds disagree"
        blob[f"{k}/found"] = np.array(ok); blob[f"{k}/true"] = uv
        blob[f"{k}/grid"] = grid.astype(np.int32); blob[f"{k}/refined"] = refined
        blob[f"{k}/start"] = start if ok else np.zeros((54, 5)); blob[f"{k}/iters"] = iters if ok else np.zeros(54, np.int32)
        for s, sigma in enumerate(SIGMAS):
            st = ref.stages(img, sigma)
            blob[f"{k}/scale{s}/cand"] = st["cand"]; blob[f"{k}/scale{s}/avg"] = np.array(st["avg"])
            blob[f"{k}/scale{s}/pattern"] = st["pattern"]
        print(k, seed, variant, "found", ok, "refined-vs-true max %.3f px" % (np.abs(refined - uv).max() if ok else np.nan),
              "scales", [len(blob[f"{k}/scale{s}/pattern"]) == 54 for s in range(3)])
    blob["images"] = np.stack(imgs)
    # SubpixelCorner::Evaluate on image 0's maps (scale 1.4) at perturbed start values
    m = ref.maps(imgs[0], 1.4)
    pr, xs, ln, cost, grad = [], [], [], [], []
    for i in range(0, 54, 3):
        prior = blob["0/grid"][i].astype(np.float64)
        for _ in range(2):
            x = blob["0/start"][i] + rng.normal(0, [0.4, 0.4, 0.05, 0.05, 0.3])
            length = float(rng.uniform(2.0, 7.0))
            c, g = ref.subpixel_evaluate(m["gradx"], m["grady"], prior, length, x)
            pr.append(prior); xs.append(x); ln.append(length); cost.append(c); grad.append(g)
    # a corner near the image border: the grid clamps the bicubic neighbourhood
    for prior, x in (((2.0, 3.0), (1.2, 2.1, 0.3, 1.9, 0.2)), ((W - 2.0, H - 1.0), (W - 1.5, H - 1.2, -0.4, 1.2, -0.3))):
        c, g = ref.subpixel_evaluate(m["gradx"], m["grady"], prior, 7.0, x)
        pr.append(np.array(prior)); xs.append(np.array(x)); ln.append(7.0); cost.append(c); grad.append(g)
    blob["eval/prior"] = np.array(pr); blob["eval/x"] = np.array(xs); blob["eval/length"] = np.array(ln)
    blob["eval/cost"] = np.array(cost); blob["eval/grad"] = np.array(grad)
    path = os.path.join(HERE, "detector.npz")
    np.savez_compressed(path, **blob)
    print(f"wrote {path}: {len(blob)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()

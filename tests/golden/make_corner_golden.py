"""Golden vectors for the detector's blur stage, from OpenCV itself (cv2 python wheel of the authoring container; the
reference calls cv::GaussianBlur, corner_detector.cpp:266,270, and OpenCV is not under /root/reference):
    python tests/golden/make_corner_golden.py
Renders synthetic board images (synthdata.render_board_image), blurs them with cv2.GaussianBlur for the kernel sizes /
sigmas computeResponse uses (3 / 0.7 and 1 + 2 ceil(s) / s for s in 1.4, 2, 1: corner_detector.cpp:229-233,265,269) and
stores images + blurred images -> corner_response.npz.  The oracle (oracle/corner_oracle.c) must reproduce the blurs bit
for bit (tests/test_corner.py); the stencil after the blur has no third-party arithmetic in it."""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import synthdata as sd  # noqa: E402

KERNELS = ((3, 0.7), (5, 1.4), (5, 2.0), (3, 1.0))


def main():
    blob = {"cv2_version": np.array(cv2.__version__)}
    for name, (w, h, seed, model) in {"eucm_173x131": (173, 131, 20250, sd.EUCM), "mei_160x120": (160, 120, 20251, sd.MEI),
                                      "ucm_64x96": (64, 96, 20252, sd.UCM)}.items():
        img, uv = sd.render_board_image(w, h, seed=seed, model=model)
        blob[f"{name}/img"] = img; blob[f"{name}/corners"] = uv
        for n, s in KERNELS:
            blob[f"{name}/blur_{n}_{s}"] = cv2.GaussianBlur(img, (n, n), s, sigmaY=s)
    rng = np.random.default_rng(7)
    noise = rng.integers(0, 256, (37, 45), dtype=np.uint8)        # worst case for rounding: white noise, tiny image
    blob["noise_45x37/img"] = noise
    for n, s in KERNELS:
        blob[f"noise_45x37/blur_{n}_{s}"] = cv2.GaussianBlur(noise, (n, n), s, sigmaY=s)
    path = os.path.join(HERE, "corner_response.npz")
    np.savez_compressed(path, **blob)
    print(f"wrote {path}: {len(blob)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()

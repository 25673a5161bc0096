"""Timing of the checkerboard detector (vg_detect_pattern, host buffers in, corners out) next to the reference's own
corner_detector.cpp (oracle/_ref, one image per host thread) on the same rendered 1280 x 800 boards.
    python tools/detector_timing.py [n_images]"""
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import synthdata as sd  # noqa: E402
import visgeom_b200 as vg  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    base = []
    for k in range(8):
        img, uv = sd.render_board_image(1280, 800, seed=20300 + k, model=(sd.EUCM, sd.MEI, sd.UCM)[k % 3], supersample=2)
        base.append((img, uv))
    imgs = np.stack([base[k % 8][0] for k in range(n)])
    truth = np.stack([base[k % 8][1] for k in range(n)])
    vg.detect_pattern(imgs[:2])                                  # context, allocations
    os.environ["VG_DETECT_TRACE"] = "1"
    for improve in (False, True):
        t = []
        for _ in range(3):
            t0 = time.perf_counter()
            found, c = vg.detect_pattern(imgs, improve=improve)
            t.append(time.perf_counter() - t0)
        dt = min(t)
        err = np.nanmax(np.abs(c - truth))
        print(f"vg_detect_pattern improve={int(improve)}: {n} images 1280x800 in {dt * 1e3:.1f} ms -> {n / dt:.0f} images/s, "
              f"{n * 1280 * 800 / dt / 1e9:.2f} Gpixel/s; found {int(found.sum())}/{n}; max |corner - truth| {err:.3f} px")
    try:
        from oracle.pyoracle import ReferenceDetector
        ref = ReferenceDetector()
    except Exception as e:
        print(f"reference build not available: {e}")
        return
    m = min(n, 16)
    t0 = time.perf_counter()
    for k in range(m):
        ref.detect_pattern(imgs[k], improve=True)
    dt1 = time.perf_counter() - t0
    print(f"reference corner_detector.cpp (oracle/_ref, -O2), 1 thread: {m / dt1:.1f} images/s ({dt1 / m * 1e3:.1f} ms per image)")
    cores = len(os.sched_getaffinity(0))
    with ThreadPoolExecutor(cores) as ex:                        # ctypes releases the GIL
        t0 = time.perf_counter()
        list(ex.map(lambda k: ref.detect_pattern(imgs[k % n], improve=True), range(4 * cores)))
        dtn = time.perf_counter() - t0
    print(f"reference corner_detector.cpp, {cores} threads (one image each): {4 * cores / dtn:.1f} images/s")


if __name__ == "__main__":
    main()

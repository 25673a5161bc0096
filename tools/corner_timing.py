"""Developer timing of the detector's response kernel (device-resident batch): algorithmic bytes = 17 B per pixel
(1 B image read, four float maps written)."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synthdata as sd
import visgeom_b200 as vg


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    w, h = 1280, 800
    dev = torch.device("cuda:0")
    base = np.stack([sd.render_board_image(w, h, seed=400 + i, supersample=1)[0] for i in range(4)])
    imgs = torch.from_numpy(np.concatenate([base] * (n // 4))).to(dev)
    outs = [torch.empty(n, h, w, dtype=torch.float32, device=dev) for _ in range(4)]
    avg = torch.empty(n, dtype=torch.float64, device=dev)
    cnt = torch.empty(n, dtype=torch.int64, device=dev)
    L = vg.lib()
    st = torch.cuda.current_stream().cuda_stream

    def run():
        rc = L.vg_corner_response_dev(imgs.data_ptr(), n, w, h, 0.7, 1.4, *[o.data_ptr() for o in outs], avg.data_ptr(), cnt.data_ptr(), st)
        assert rc == 0
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    by = 17.0 * n * w * h
    peak = 6458.4
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    print(f"corner response: {n} images {w}x{h}: {us:.1f} us per batch, {n * w * h / us / 1e3:.2f} Gpixel/s, "
          f"{by / us / 1e3:.0f} GB/s algorithmic = {by / us / 1e3 / peak:.3f} of the measured HBM peak ({peak:.0f} GB/s)")


if __name__ == "__main__":
    main()

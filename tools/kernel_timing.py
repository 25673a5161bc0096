"""Developer timing of the fused kernel in its output modes (not the bench contract)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synthdata as sd
import visgeom_b200 as vg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", type=int, default=0)
    ap.add_argument("--n-img", type=int, default=10000)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--sets", type=int, default=4)
    ap.add_argument("--modes", default="full,jac,normal,resid")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    d = sd.make_mono(a.model, a.n_img, seed=20242)
    K, P, n = d["K"], d["P"], a.n_img
    ne = vg.hessian_entries(a.model, 1)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    intr, board, xi = t(d["intr_init"]), t(d["board"]), t(d["xi_init"])
    obs = [t(d["obs"]) for _ in range(a.sets)]
    r = [torch.empty(n, 2 * P, dtype=torch.float64, device=dev) for _ in range(a.sets)]
    Ja = [torch.empty(n, 2 * P, K, dtype=torch.float64, device=dev) for _ in range(a.sets)]
    Je = [torch.empty(n, 2 * P, 6, dtype=torch.float64, device=dev) for _ in range(a.sets)]
    H = [torch.empty(n, ne, dtype=torch.float64, device=dev) for _ in range(a.sets)]
    stream = torch.cuda.current_stream().cuda_stream
    bytes_full = n * (P * (16 + 16 + 96 + 16 * K) + 48 + ne * 8)

    def run(mode, s):
        vg.eval_chain_dev(a.model, intr.data_ptr(), board.data_ptr(), obs[s].data_ptr(), [xi.data_ptr()], [0], [0],
                          n, P,
                          r=r[s].data_ptr() if mode != "normal" else None,
                          J_intr=Ja[s].data_ptr() if mode in ("full", "jac") else None,
                          J_xi=[Je[s].data_ptr()] if mode in ("full", "jac") else None,
                          H=H[s].data_ptr() if mode in ("full", "normal") else None, stream=stream)

    for mode in a.modes.split(","):
        for i in range(10):
            run(mode, i % a.sets)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.steps):
            run(mode, i % a.sets)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / a.steps
        by = {"full": bytes_full, "jac": bytes_full - n * ne * 8,
              "normal": n * (P * 16 + 48 + ne * 8), "resid": n * (P * 32 + 48)}[mode]
        print(f"mode={mode:7s} {us:9.2f} us/launch  {n * P / us:10.1f} Mcorner/s  {by / us / 1e3:8.1f} GB/s algorithmic")


if __name__ == "__main__":
    main()

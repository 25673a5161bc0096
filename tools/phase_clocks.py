"""Developer tool: per-phase clock64 breakdown of the evaluation kernel.
Run with VG_VARIANT=phase (build: VG_VARIANT=phase python -m visgeom_b200.build)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synthdata as sd
import visgeom_b200 as vg

PHASED = True
NAMES = [["A corner", "H wait+fence", "B2 barrier", "S/B tma,gram", "B3 barrier", "H out", "-", "-"]] * 2


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    mode = sys.argv[2] if len(sys.argv) > 2 else "full"
    dev = torch.device("cuda:0")
    d = sd.make_mono(0, n, seed=20242)
    K, P = d["K"], d["P"]
    ne = vg.hessian_entries(0, 1)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    intr, board, xi, obs = t(d["intr_init"]), t(d["board"]), t(d["xi_init"]), t(d["obs"])
    r = torch.empty(n, 2 * P, dtype=torch.float64, device=dev)
    Ja = torch.empty(n, 2 * P, K, dtype=torch.float64, device=dev)
    Je = torch.empty(n, 2 * P, 6, dtype=torch.float64, device=dev)
    H = torch.empty(n, ne, dtype=torch.float64, device=dev)
    L = vg.lib()
    out = (C.c_ulonglong * 16)()
    reps = 20

    def run():
        vg.eval_chain_dev(0, intr.data_ptr(), board.data_ptr(), obs.data_ptr(), [xi.data_ptr()], [0], [0], n, P,
                          r=r.data_ptr() if mode != "normal" else None,
                          J_intr=Ja.data_ptr() if mode == "full" else None,
                          J_xi=[Je.data_ptr()] if mode == "full" else None,
                          H=H.data_ptr() if mode in ("full", "normal") else None,
                          stream=torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        run()
    L.vg_debug_phase_clocks(out, 1)
    for _ in range(reps):
        run()
    L.vg_debug_phase_clocks(out, 1)
    v = np.array(list(out), dtype=np.float64).reshape(2, 8) / reps
    grid = 592 if PHASED else 148
    for w, nm in ((0, "warp 0 (Gram)"), (1, "last warp (TMA issue, poses)")):
        tot = v[w].sum()
        print(f"{nm}: {tot / grid:9.0f} cycles per CTA")
        for i in range(8):
            print(f"   {NAMES[w][i]:14s} {v[w, i] / grid:9.0f} cycles/CTA  {100 * v[w, i] / tot:5.1f}%")


if __name__ == "__main__":
    main()

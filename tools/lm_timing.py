"""Developer timing of vg_problem_solve on the C2 problem (VG_LM_TRACE=1 prints the per-iteration device times)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synthdata as sd
import visgeom_b200 as vg


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    model = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    d = sd.make_mono(model, n, seed=20242)
    for rep in range(3):
        Pm = vg.Problem(0)
        cam = Pm.add_camera(model, d["intr_init"])
        tr = Pm.add_transform(d["xi_init"], is_global=False)
        Pm.add_dataset(cam, d["board"], d["obs"], [tr], [0])
        o = Pm.default_options()
        o.max_num_iterations = 25
        Pm.evaluate()
        t0 = time.perf_counter()
        sm = Pm.solve(o)
        dt = time.perf_counter() - t0
        print(f"iterations {sm.iterations} in {dt * 1e3:.3f} ms -> {sm.iterations / dt:.0f} iterations/s; "
              f"{dt / sm.iterations * 1e6:.1f} us per iteration; final cost {sm.final_cost:.6f}", flush=True)
        Pm.close()


if __name__ == "__main__":
    main()

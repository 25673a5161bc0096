"""Developer tool: one vg_problem_solve of the C2 problem; prints iterations/s (run under ncu for the launch list)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synthdata as sd
import visgeom_b200 as vg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 25
d = sd.make_mono(sd.EUCM, n, seed=20242)
for rep in range(2):
    P = vg.Problem(0)
    cam = P.add_camera(sd.EUCM, d["intr_init"])
    tr = P.add_transform(d["xi_init"], is_global=False)
    P.add_dataset(cam, d["board"], d["obs"], [tr], [0])
    P.evaluate()
    o = P.default_options()
    o.max_num_iterations = iters
    t0 = time.perf_counter()
    s = P.solve(o)
    dt = time.perf_counter() - t0
    print(f"iterations {s.iterations} in {dt * 1e3:.3f} ms -> {s.iterations / dt:.0f} iterations/s; "
          f"{dt / max(1, s.iterations) * 1e6:.1f} us per iteration; evaluate {s.seconds_evaluate * 1e3:.3f} ms over {s.num_evaluations}")

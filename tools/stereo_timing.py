"""Developer tool: config C4 (stereo EUCM, 5 000 pairs: camera1 chain [board], camera2 chain [xiCam12 inverse, board])
-- time per evaluation of the whole problem (two kernel launches: L = 1 and L = 2) and LM iterations/s."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synthdata as sd
import visgeom_b200 as vg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
s = sd.make_stereo(n, seed=20244)


def build():
    P = vg.Problem(0)
    c1 = P.add_camera(sd.EUCM, s["intr1_init"]); c2 = P.add_camera(sd.EUCM, s["intr2_init"])
    t12 = P.add_transform(s["xi12_init"], is_global=True)
    tb = P.add_transform(s["xi_init"], is_global=False)
    P.add_dataset(c1, s["board"], s["obs1"], [tb], [0])
    P.add_dataset(c2, s["board"], s["obs2"], [t12, tb], [1, 0])
    return P


P = build()
P.evaluate()
for _ in range(20):
    P.evaluate_async()
P.fetch_reduced()
reps = 300
t0 = time.perf_counter()
for _ in range(reps):
    P.evaluate_async()
P.fetch_reduced()
dt = (time.perf_counter() - t0) / reps
# algorithmic bytes per pair, SURVEY 8d: 30 632 (normal-equation blocks only when the Jacobians stay on chip)
print(f"C4 evaluation (normal equations only): {dt * 1e6:.1f} us for {2 * n * 54} corners -> {2 * n * 54 / dt / 1e9:.2f} G corner evaluations/s")
for rep in range(2):
    P = build()
    P.evaluate()
    o = P.default_options(); o.max_num_iterations = 25
    t0 = time.perf_counter()
    sm = P.solve(o)
    dt = time.perf_counter() - t0
    print(f"C4 LM: {sm.iterations} iterations in {dt * 1e3:.3f} ms -> {sm.iterations / dt:.0f} iterations/s; final cost {sm.final_cost:.6f}")

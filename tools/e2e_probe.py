"""Developer probe: the end-to-end leg of bench.py (host inputs every step) with 1 / 2 / 3 problems in flight, and its parts."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import synthdata as sd
import visgeom_b200 as vg

n_img, P = 10000, 54
d = sd.make_mono(sd.EUCM, n_img, seed=20242)
h_obs = torch.from_numpy(d["obs"]).pin_memory()
h_xi = torch.from_numpy(d["xi_init"]).pin_memory()


def make():
    Pm = vg.Problem(0)
    cam = Pm.add_camera(sd.EUCM, d["intr_init"])
    tr = Pm.add_transform(d["xi_init"], is_global=False)
    ds = Pm.add_dataset(cam, d["board"], d["obs"], [tr], [0])
    Pm.evaluate()
    return Pm, cam, tr, ds


def run(depth, what, n=200):
    probs = [make() for _ in range(depth)]

    def issue(i):
        Pm, cam, tr, ds = probs[i % depth]
        if "obs" in what: Pm.update_observations(ds, h_obs.data_ptr())
        if "xi" in what: Pm.update_poses(tr, h_xi.data_ptr())
        if "cam" in what: Pm.set_camera(cam, d["intr_init"])
        if "eval" in what: Pm.evaluate_async()

    def fetch(i):
        if "fetch" in what:
            return probs[i % depth][0].fetch_reduced()
    for i in range(2 * depth):
        issue(i); fetch(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(depth - 1):
        issue(i)
    for i in range(depth - 1, n):
        issue(i)
        fetch(i - (depth - 1))
    for i in range(n - (depth - 1), n):
        fetch(i)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    print(f"depth {depth} {'+'.join(what):28s}: {dt * 1e6:7.1f} us per step -> {n_img * P / dt / 1e9:.2f} G corner evaluations/s", flush=True)
    for p in probs:
        p[0].close()


full = ["obs", "xi", "cam", "eval", "fetch"]
for depth in (1, 2, 3, 4):
    run(depth, full)
run(2, ["obs", "eval", "fetch"])
run(2, ["obs", "xi", "eval", "fetch"])
run(2, ["obs"])
run(2, ["eval", "fetch"])
run(2, ["obs", "xi", "cam", "eval"])

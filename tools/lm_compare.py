import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, synthdata as sd
n = int(sys.argv[1]); which = sys.argv[2]
d = sd.make_mono(0, n, seed=20242)
if which == "oracle":
    from oracle.pyoracle import Oracle, OracleProblem
    Pm = OracleProblem(Oracle())
else:
    import visgeom_b200 as vg
    Pm = vg.Problem(0)
cam = Pm.add_camera(0, d["intr_init"]); tr = Pm.add_transform(d["xi_init"], is_global=False)
Pm.add_dataset(cam, d["board"], d["obs"], [tr], [0])
o = Pm.default_options() if hasattr(Pm, "default_options") else Pm.o.default_options()
o.max_num_iterations = 25; o.verbose = 1
sm = Pm.solve(o)
print(which, "iterations", sm.iterations, "succ", sm.num_successful, "unsucc", sm.num_unsuccessful, "term", sm.termination, "cost", sm.final_cost, Pm.camera(cam))

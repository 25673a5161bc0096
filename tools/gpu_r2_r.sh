#!/bin/bash
# round 2, GPU call R: camera point API, chunked per-thread vg_eval_chain, full suite, bench
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py --steps 100 --warmup 10 --cpu-seconds 2 > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo "bench rc $?"; tail -c 300 gpurun_out/r2r_bench.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/r2r_bench.json").read().strip().splitlines()[-1])
print({k:l[k] for k in ("value","ms_per_step")}, l["lm"]["iters_per_s"], l["roofline"]["frac"], l["roofline"]["step_frac"], "e2e", l["e2e"]["value"], "ceres", l["e2e_ceres_contract"]["value"])
PY
for t in 1 2 4 8; do for mb in 4 8 16; do echo "threads $t chunk $mb MB"; VG_HOST_COPY_THREADS=$t VG_HOST_CHUNK_MB=$mb python - <<'PY'
import time, sys
sys.path.insert(0, ".")
import synthdata as sd, visgeom_b200 as vg
d = sd.make_mono(0, 10000, seed=20242)
a = (0, d["intr_init"], d["board"], d["obs"], [d["xi_init"]], [0], [0])
o = vg.eval_chain(*a, want_H=True)
vg.eval_chain(*a, want_H=True, out=o)
t0 = time.perf_counter()
for _ in range(5): vg.eval_chain(*a, want_H=True, out=o)
dt = (time.perf_counter() - t0) / 5
print(f"  {dt*1e3:.2f} ms per call -> {540000/dt/1e6:.1f} M corner evaluations/s, {127.76e6/dt/1e9:.1f} GB/s back")
PY
done; done

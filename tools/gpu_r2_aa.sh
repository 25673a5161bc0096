#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/kernel_timing.py --n-img 10000 --steps 300 --modes full,normal
python tools/stereo_timing.py 2>&1 | tail -3
for k in 1 2; do timeout 900 python bench.py --steps 20 --warmup 5 --cpu-seconds 1 > gpurun_out/r2aa_bench_$k.json 2>/dev/null; python - <<PY
import json
l=json.loads(open("gpurun_out/r2aa_bench_$k.json").read().strip().splitlines()[-1])
print(l["ms_per_step"], l["roofline"]["kernel_us"], l["roofline"]["frac"], l["roofline"]["step_frac"], l["lm"]["iters_per_s"], l["e2e"]["value"], l["e2e_ceres_contract"]["value"])
PY
done

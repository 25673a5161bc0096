#!/bin/bash
# round 2, GPU call H: sliced reduction tail -> step time; solve tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_solve_gpu.py tests/test_eval_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 900 python bench.py --steps 200 --warmup 20 --cpu-seconds 2 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc $?"; tail -c 300 gpurun_out/r2h_bench.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/r2h_bench.json").read().strip().splitlines()[-1])
print({k:l[k] for k in ("value","ms_per_step","e2e","lm")}); print(l["roofline"])
PY

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 100 --warmup 10 --cpu-seconds 2 > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc $?"; tail -c 300 gpurun_out/r2p_bench.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/r2p_bench.json").read().strip().splitlines()[-1])
print({k:l[k] for k in ("value","ms_per_step")}, l["lm"]["iters_per_s"], l["lm"]["iterations"], l["roofline"]["frac"], l["roofline"]["step_frac"], l["e2e"]["value"], l["e2e_ceres_contract"]["value"])
PY

#!/bin/bash
# round 2, GPU call I: direct assembly in the reduction tail (step time), batched finalize loads in the solver kernels (LM)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
{ python tools/lm_timing.py 10000 0; VG_LM_TRACE=1 python tools/lm_timing.py 10000 0 2>&1 | tail -12; } > gpurun_out/r2i_lm.txt 2>&1; cat gpurun_out/r2i_lm.txt
timeout 900 python bench.py --steps 200 --warmup 20 --cpu-seconds 2 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc $?"; tail -c 300 gpurun_out/r2i_bench.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/r2i_bench.json").read().strip().splitlines()[-1])
print({k:l[k] for k in ("value","ms_per_step")}, l["lm"]["iters_per_s"], l["roofline"]["frac"], l["roofline"]["step_frac"], l["roofline"]["kernel_us"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2i_lm_launches.csv python tools/lm_timing.py 10000 0 > /dev/null 2>&1

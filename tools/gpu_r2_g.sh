#!/bin/bash
# round 2, GPU call G: cleaned-up kernel (two barriers per group), parity-Gram variant, phase clocks, new bench.py
mkdir -p gpurun_out
{
for v in "" parity; do
  echo "=== variant '$v'"
  VG_VARIANT=$v timeout 300 python -m pytest tests/test_eval_gpu.py -m gpu -x -q 2>&1 | tail -1
  VG_VARIANT=$v timeout 300 python tools/kernel_timing.py --n-img 10000 --steps 300 --modes full,normal,resid
  VG_VARIANT=$v timeout 300 python tools/kernel_timing.py --n-img 25000 --steps 200 --modes full
  VG_VARIANT=$v timeout 300 python tools/kernel_timing.py --n-img 10000 --model 2 --steps 200 --modes full
  VG_VARIANT=$v timeout 300 python tools/kernel_timing.py --n-img 10000 --model 1 --steps 200 --modes full
done
echo "== phase clocks"; VG_VARIANT=phase timeout 300 python tools/phase_clocks.py 10000 full
} > gpurun_out/r2g_timing.txt 2>&1
cat gpurun_out/r2g_timing.txt
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc $?"; tail -c 300 gpurun_out/r2g_bench.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/r2g_bench.json").read().strip().splitlines()[-1])
print({k:l[k] for k in ("value","ms_per_step","roofline","e2e","e2e_ceres_contract","lm","config")})
PY

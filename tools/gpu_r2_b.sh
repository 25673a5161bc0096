#!/bin/bash
# round 2, GPU call B: A/B of the kernel's restructuring switches (each variant: parity subset + timing)
mkdir -p gpurun_out
{
for v in r1 "" alloff nosplit noswap strided camreg; do
  echo "=== variant '$v'"
  if [ "$v" != "r1" ]; then VG_VARIANT=$v timeout 300 python -m pytest tests/test_eval_gpu.py -m gpu -x -q 2>&1 | tail -1; fi
  VG_VARIANT=$v timeout 300 python tools/kernel_timing.py --n-img 10000 --steps 300 --modes full,normal,resid
  VG_VARIANT=$v timeout 300 python tools/kernel_timing.py --n-img 25000 --steps 200 --modes full
  VG_VARIANT=$v timeout 300 python tools/kernel_timing.py --n-img 10000 --model 1 --steps 200 --modes full
done
} > gpurun_out/r2b_timing.txt 2>&1
cat gpurun_out/r2b_timing.txt

#!/bin/bash
# round 2, GPU call C: bisect of the loop restructuring (b*: one structural change each on the round-1 shape; c*: one data-path change each)
mkdir -p gpurun_out
{
for v in r1 b0 b1 b2 b3 c0 c1 c2 c3 ""; do
  echo "=== variant '$v'"
  if [ "$v" != "r1" ]; then VG_VARIANT=$v timeout 300 python -m pytest tests/test_eval_gpu.py -m gpu -x -q 2>&1 | tail -1; fi
  VG_VARIANT=$v timeout 300 python tools/kernel_timing.py --n-img 10000 --steps 300 --modes full,resid
  VG_VARIANT=$v timeout 300 python tools/kernel_timing.py --n-img 25000 --steps 200 --modes full
done
} > gpurun_out/r2c_timing.txt 2>&1
cat gpurun_out/r2c_timing.txt

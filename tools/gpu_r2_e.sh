#!/bin/bash
# round 2, GPU call E: cumulative ladder n0 (round-1 shape in the new skeleton) .. n6, default = everything on
mkdir -p gpurun_out
{
for v in r1 n0 n1 n2 n3 n4 n5 n6 ""; do
  echo "=== variant '$v'"
  if [ "$v" != "r1" ]; then VG_VARIANT=$v timeout 300 python -m pytest tests/test_eval_gpu.py -m gpu -x -q 2>&1 | tail -1; fi
  VG_VARIANT=$v timeout 300 python tools/kernel_timing.py --n-img 10000 --steps 300 --modes full,resid
  VG_VARIANT=$v timeout 300 python tools/kernel_timing.py --n-img 25000 --steps 200 --modes full
done
} > gpurun_out/r2e_timing.txt 2>&1
cat gpurun_out/r2e_timing.txt

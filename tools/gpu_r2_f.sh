#!/bin/bash
# round 2, GPU call F: board compiled in (default) vs run-time sizes (nofix), + restage / split on top; full parity run of the default
mkdir -p gpurun_out
{
for v in "" nofix restage split; do
  echo "=== variant '$v'"
  VG_VARIANT=$v timeout 300 python -m pytest tests/test_eval_gpu.py -m gpu -x -q 2>&1 | tail -1
  VG_VARIANT=$v timeout 300 python tools/kernel_timing.py --n-img 10000 --steps 300 --modes full,normal,resid
  VG_VARIANT=$v timeout 300 python tools/kernel_timing.py --n-img 25000 --steps 200 --modes full
  VG_VARIANT=$v timeout 300 python tools/kernel_timing.py --n-img 10000 --model 2 --steps 200 --modes full
  VG_VARIANT=$v timeout 300 python tools/kernel_timing.py --n-img 10000 --model 1 --steps 200 --modes full
done
} > gpurun_out/r2f_timing.txt 2>&1
cat gpurun_out/r2f_timing.txt
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:reproj -s 12 -c 1 -o gpurun_out/r2f_new -f python tools/kernel_timing.py --modes full --steps 20 --n-img 10000 > gpurun_out/r2f_ncu.log 2>&1
ncu -i gpurun_out/r2f_new.ncu-rep --page raw --csv > gpurun_out/r2f_new_raw.csv 2>/dev/null
ncu -i gpurun_out/r2f_new.ncu-rep --page source --csv --print-source sass > gpurun_out/r2f_new_sass.csv 2>/dev/null

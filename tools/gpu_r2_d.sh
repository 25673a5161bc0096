#!/bin/bash
# round 2, GPU call D: ncu of the round-1 kernel vs the restructured ones (why is the new skeleton slower?)
mkdir -p gpurun_out
for v in r1 b0 b1 ""; do
  n=${v:-new}
  VG_VARIANT=$v timeout 600 ncu --set full --clock-control none --import-source on -k regex:reproj -s 12 -c 1 -o gpurun_out/r2d_$n -f python tools/kernel_timing.py --modes full --steps 20 --n-img 10000 > gpurun_out/r2d_$n.log 2>&1
  ncu -i gpurun_out/r2d_$n.ncu-rep --page raw --csv > gpurun_out/r2d_${n}_raw.csv 2>/dev/null
done
ls -la gpurun_out/r2d_*

#!/bin/bash
mkdir -p gpurun_out
for k in fast_backsub; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -o gpurun_out/r2o_$k -f python tools/lm_timing.py 10000 0 > gpurun_out/r2o_$k.log 2>&1
ncu -i gpurun_out/r2o_$k.ncu-rep --page raw --csv > gpurun_out/r2o_${k}_raw.csv 2>/dev/null
ncu -i gpurun_out/r2o_$k.ncu-rep --page source --csv > gpurun_out/r2o_${k}_src.csv 2>/dev/null
done

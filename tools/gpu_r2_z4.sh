#!/bin/bash
# ncu of fast_factor / fast_backsub inside the device-driven LM loop
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fast_ -s 4 -c 2 -o gpurun_out/r2z4 -f python tools/lm_timing.py 10000 0 > gpurun_out/r2z4.log 2>&1
ncu -i gpurun_out/r2z4.ncu-rep --page raw --csv > gpurun_out/r2z4_raw.csv 2>/dev/null
ncu -i gpurun_out/r2z4.ncu-rep --page source --csv --print-source sass > gpurun_out/r2z4_sass.csv 2>/dev/null
rm -f gpurun_out/r2z4.ncu-rep
tail -3 gpurun_out/r2z4.log

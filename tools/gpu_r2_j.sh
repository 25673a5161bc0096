#!/bin/bash
# round 2, GPU call J: ncu of pose_factor / pose_backsub inside the LM loop
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pose_factor -s 6 -c 1 -o gpurun_out/r2j_factor -f python tools/lm_timing.py 10000 0 > gpurun_out/r2j_factor.log 2>&1
ncu -i gpurun_out/r2j_factor.ncu-rep --page raw --csv > gpurun_out/r2j_factor_raw.csv 2>/dev/null
ncu -i gpurun_out/r2j_factor.ncu-rep --page source --csv --print-source sass > gpurun_out/r2j_factor_sass.csv 2>/dev/null
ncu -i gpurun_out/r2j_factor.ncu-rep --page source --csv > gpurun_out/r2j_factor_src.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pose_backsub -s 6 -c 1 -o gpurun_out/r2j_backsub -f python tools/lm_timing.py 10000 0 > gpurun_out/r2j_backsub.log 2>&1
ncu -i gpurun_out/r2j_backsub.ncu-rep --page raw --csv > gpurun_out/r2j_backsub_raw.csv 2>/dev/null
ncu -i gpurun_out/r2j_backsub.ncu-rep --page source --csv > gpurun_out/r2j_backsub_src.csv 2>/dev/null
ls -la gpurun_out/r2j_*

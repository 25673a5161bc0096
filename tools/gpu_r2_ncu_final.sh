#!/bin/bash
# ncu --set full of the evaluation kernel at C2 with the final code of the round (a lone launch: ncu serialises)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:reproj -s 12 -c 1 -o gpurun_out/fk -f \
    python tools/kernel_timing.py --modes full --steps 20 --n-img 10000 --model 0 > gpurun_out/fk.log 2>&1
ncu -i gpurun_out/fk.ncu-rep --page raw --csv > gpurun_out/fk_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/fk_raw.csv > gpurun_out/fk_summary.txt
rm -f gpurun_out/fk.ncu-rep
head -8 gpurun_out/fk_summary.txt

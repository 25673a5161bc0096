#!/bin/bash
# ncu --set full of the evaluation kernel at the C5 shard size and for MEI with the final code of the round (lone launches)
mkdir -p gpurun_out
for cfg in "25k 25000 0" "mei 10000 2"; do
  set -- $cfg
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:reproj -s 12 -c 1 -o gpurun_out/fk_$1 -f \
      python tools/kernel_timing.py --modes full --steps 20 --n-img $2 --model $3 > gpurun_out/fk_$1.log 2>&1
  ncu -i gpurun_out/fk_$1.ncu-rep --page raw --csv > gpurun_out/fk_$1_raw.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/fk_$1_raw.csv > gpurun_out/fk_$1_summary.txt
  rm -f gpurun_out/fk_$1.ncu-rep gpurun_out/fk_$1_raw.csv
  head -3 gpurun_out/fk_$1_summary.txt
done

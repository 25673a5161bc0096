#!/bin/bash
# detector: timing with the stage trace, then ncu --set full of its three kernels
mkdir -p gpurun_out
timeout 600 python tools/detector_timing.py 64 > gpurun_out/det3_timing.log 2>&1
cat gpurun_out/det3_timing.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"local_maxima|subpixel_refine|corner_response" -s 3 -c 6 \
    -o gpurun_out/det3 -f python tools/detector_timing.py 16 > gpurun_out/det3_ncu.log 2>&1
ncu -i gpurun_out/det3.ncu-rep --page raw --csv > gpurun_out/det3_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/det3_raw.csv --all > gpurun_out/det3_summary.txt
rm -f gpurun_out/det3.ncu-rep
grep -c . gpurun_out/det3_summary.txt

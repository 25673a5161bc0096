#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_corner.py tests/test_camera_api.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/corner_timing.py 64 2>&1 | tee gpurun_out/r2s_corner.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:corner_response -s 3 -c 1 -o gpurun_out/r2s_corner -f python tools/corner_timing.py 64 > gpurun_out/r2s_corner_ncu.log 2>&1
ncu -i gpurun_out/r2s_corner.ncu-rep --page raw --csv > gpurun_out/r2s_corner_raw.csv 2>/dev/null

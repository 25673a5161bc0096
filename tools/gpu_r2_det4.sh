#!/bin/bash
# detector: the random-picture parity test, then ncu --set full of the refinement kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_detector_gpu.py -x -q -m gpu > gpurun_out/det4_tests.log 2>&1
tail -6 gpurun_out/det4_tests.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:subpixel_refine -s 1 -c 1 \
    -o gpurun_out/det4 -f python tools/detector_timing.py 16 > gpurun_out/det4_ncu.log 2>&1
ncu -i gpurun_out/det4.ncu-rep --page raw --csv > gpurun_out/det4_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/det4_raw.csv --all > gpurun_out/det4_summary.txt
rm -f gpurun_out/det4.ncu-rep
head -20 gpurun_out/det4_summary.txt
python tools/detector_timing.py 64 2>&1 | grep "improve=1"

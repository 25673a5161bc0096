#!/bin/bash
# detector: GPU parity tests + corner tests (kernel changed) + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_detector_gpu.py tests/test_corner.py -x -q -m gpu > gpurun_out/det1_tests.log 2>&1
tail -15 gpurun_out/det1_tests.log
timeout 600 python tools/detector_timing.py 64 > gpurun_out/det1_timing.log 2>&1
cat gpurun_out/det1_timing.log

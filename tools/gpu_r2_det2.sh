#!/bin/bash
# detector: GPU parity tests, corner tests (kernel changed), the "images" dataset through vg_calib, timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_detector_gpu.py tests/test_corner.py tests/test_calib_cli.py -x -q -m gpu > gpurun_out/det2_tests.log 2>&1
tail -15 gpurun_out/det2_tests.log
timeout 600 python tools/detector_timing.py 64 > gpurun_out/det2_timing.log 2>&1
cat gpurun_out/det2_timing.log

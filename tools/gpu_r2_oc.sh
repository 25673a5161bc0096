#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_priors_gpu.py -x -q -m gpu > gpurun_out/oc_tests.log 2>&1
tail -12 gpurun_out/oc_tests.log
for c in 4 8 16; do echo "== VG_DETECT_CHUNK=$c"; VG_DETECT_CHUNK=$c python tools/detector_timing.py 64 2>&1 | grep "improve=1\|wall" | tail -2; done

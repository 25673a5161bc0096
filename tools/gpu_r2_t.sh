#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/corner_timing.py 64 2>&1 | tee gpurun_out/r2t_corner.txt

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_solve_gpu.py -m gpu -x -q 2>&1 | tail -2
{ python tools/lm_timing.py 10000 0; VG_LM_TRACE=1 python tools/lm_timing.py 10000 0 2>&1 | tail -7; } 2>&1 | tee gpurun_out/r2n_lm.txt

#!/bin/bash
mkdir -p gpurun_out
echo "== device loop"; python tools/lm_compare.py 10000 ours 2>&1 | tail -12
echo "== host loop"; VG_LM_HOSTLOOP=1 python tools/lm_compare.py 10000 ours 2>&1 | tail -12
echo "== timing device loop"; python tools/lm_timing.py 2>&1 | tail -3
echo "== timing host loop"; VG_LM_HOSTLOOP=1 python tools/lm_timing.py 2>&1 | tail -2
python tools/lm_timing.py 10000 2 2>&1 | tail -2
python tools/lm_timing.py 10000 1 2>&1 | tail -2
timeout 900 python -m pytest tests/test_solve_gpu.py tests/test_peer_exchange_gpu.py tests/test_calib_cli.py -m gpu -x -q 2>&1 | tail -8
python tools/kernel_timing.py --n-img 10000 --steps 300 --modes full,normal

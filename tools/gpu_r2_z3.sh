#!/bin/bash
VG_VARIANT=stamps python tools/lm_timing.py 2>&1 | tail -7
python tools/lm_timing.py 2>&1 | tail -2
python tools/lm_timing.py 10000 2 2>&1 | tail -1
python tools/lm_timing.py 10000 1 2>&1 | tail -1
python tools/lm_compare.py 10000 ours 2>&1 | tail -3
VG_LM_HOSTLOOP=1 python tools/lm_timing.py 2>&1 | tail -1
timeout 900 python -m pytest tests/test_solve_gpu.py tests/test_peer_exchange_gpu.py -m gpu -x -q 2>&1 | tail -3

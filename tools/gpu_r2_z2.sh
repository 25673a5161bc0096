#!/bin/bash
VG_LM_TIMELINE=1 python tools/lm_timing.py 2>&1 | tail -14
python tools/lm_compare.py 10000 ours 2>&1 | tail -3
timeout 900 python -m pytest tests/test_solve_gpu.py tests/test_peer_exchange_gpu.py tests/test_calib_cli.py -m gpu -x -q 2>&1 | tail -4

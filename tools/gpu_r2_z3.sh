#!/bin/bash
python tools/lm_timing.py 2>&1 | tail -2
python tools/lm_timing.py 10000 2 2>&1 | tail -1
python tools/lm_timing.py 10000 1 2>&1 | tail -1
python tools/lm_compare.py 10000 ours 2>&1 | tail -3
timeout 900 python -m pytest tests/test_solve_gpu.py tests/test_peer_exchange_gpu.py tests/test_calib_cli.py -m gpu -x -q 2>&1 | tail -3
python tools/kernel_timing.py --n-img 10000 --steps 300 --modes full
timeout 900 python bench.py --steps 20 --warmup 5 --cpu-seconds 1 2>/dev/null | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print(l['ms_per_step'], l['roofline']['kernel_us'], l['roofline']['frac'], l['roofline']['step_frac'], l['lm']['iters_per_s'], l['e2e']['value'], l['e2e_ceres_contract']['value'])"

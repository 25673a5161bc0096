set -x
python -m pytest tests/test_solve_gpu.py tests/test_priors_gpu.py -m gpu -q -x 2>&1 | tail -2
python tools/lm_timing.py 2>&1 | tail -2
VG_LM_TRACE=1 python tools/lm_timing.py 2>&1 | grep device | tail -2

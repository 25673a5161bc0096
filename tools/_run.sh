set -x
python -m pytest tests/test_solve_gpu.py tests/test_eval_gpu.py -m gpu -q --tb=short 2>&1 | tail -4

set -x
python -m pytest tests/test_priors_gpu.py -m gpu -q --tb=short 2>&1 | tail -12

set -x
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30
python tools/kernel_timing.py --modes full,normal 2>&1 | tail -2
python tools/kernel_timing.py --model 2 --modes full 2>&1 | tail -1
python tools/lm_timing.py 2>&1 | tail -1

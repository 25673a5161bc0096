set -x
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40

python -m pytest tests/test_calib_cli.py -m gpu -q --tb=short 2>&1 | tail -12

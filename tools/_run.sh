python -m pytest tests/test_calib_cli.py -m gpu -x -q 2>&1 | tail -25

python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo default; python tools/kernel_timing.py --modes full,normal 2>&1 | tail -2
python tools/kernel_timing.py --model 2 --modes full,normal 2>&1 | tail -2
echo g1; VG_VARIANT=g1 python -m pytest tests/test_eval_gpu.py -m gpu -x -q 2>&1 | tail -2
VG_VARIANT=g1 python tools/kernel_timing.py --modes full,normal 2>&1 | tail -2
VG_VARIANT=g1 python tools/kernel_timing.py --model 1 --modes full,normal 2>&1 | tail -2
VG_VARIANT=g1 python tools/kernel_timing.py --n-img 25000 --modes full 2>&1 | tail -1

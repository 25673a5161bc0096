timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 60 python tools/kernel_timing.py --modes full,jac,normal 2>&1 | tail -3
timeout 60 python tools/kernel_timing.py --n-img 25000 --modes full 2>&1 | tail -1
timeout 60 python tools/kernel_timing.py --model 2 --modes full,normal 2>&1 | tail -2
python -c "
import visgeom_b200 as vg, ctypes as C
print('smem', vg.lib().vg_version())
"

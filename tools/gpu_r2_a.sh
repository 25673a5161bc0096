#!/bin/bash
# round 2, GPU call A: parity of the restructured evaluation kernel + A/B timing against the round-1 library
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
{
for v in r1 ""; do
  echo "== variant '$v' EUCM 10k"; VG_VARIANT=$v timeout 300 python tools/kernel_timing.py --n-img 10000 --steps 300
  echo "== variant '$v' EUCM 25k"; VG_VARIANT=$v timeout 300 python tools/kernel_timing.py --n-img 25000 --steps 200 --modes full,normal
  echo "== variant '$v' MEI 10k"; VG_VARIANT=$v timeout 300 python tools/kernel_timing.py --n-img 10000 --model 2 --steps 200 --modes full,normal
  echo "== variant '$v' UCM 10k"; VG_VARIANT=$v timeout 300 python tools/kernel_timing.py --n-img 10000 --model 1 --steps 200 --modes full,normal
done
echo "== phase clocks (new)"; VG_VARIANT=phase timeout 300 python tools/phase_clocks.py 10000 full
} > gpurun_out/r2a_timing.txt 2>&1
cat gpurun_out/r2a_timing.txt
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 1500 gpurun_out/r2a_bench.json

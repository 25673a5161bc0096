#!/bin/bash
# round 2, GPU call V: the evidence kept under profiles/r02_* (bench lines, timings, ncu launch lists and full captures)
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r02_bench_line.json 2> $O/r02_bench.err; echo "bench rc $?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/r02_bench_line_reference_arm.json 2>> $O/r02_bench.err
{
echo "# tools/kernel_timing.py (CUDA events, kernel alone, 4 rotating buffer sets)"
echo "== EUCM 10 000 images (C2)"; python tools/kernel_timing.py --n-img 10000 --steps 300
echo "== EUCM 25 000 images (C5 shard)"; python tools/kernel_timing.py --n-img 25000 --steps 200 --modes full,normal
echo "== MEI 10 000 images (C3)"; python tools/kernel_timing.py --n-img 10000 --model 2 --steps 200 --modes full,normal
echo "== UCM 10 000 images"; python tools/kernel_timing.py --n-img 10000 --model 1 --steps 200 --modes full,normal
echo "# tools/lm_timing.py: vg_problem_solve of the C2 problem (EUCM), then MEI, UCM"
python tools/lm_timing.py 10000 0; python tools/lm_timing.py 10000 2; python tools/lm_timing.py 10000 1
echo "# the general kernels on the same problem (VG_LM_NOFAST=1)"
VG_LM_NOFAST=1 python tools/lm_timing.py 10000 0
echo "# per-stage device times (VG_LM_TRACE=1: events between the launches, no polling)"
VG_LM_TRACE=1 python tools/lm_timing.py 10000 0 2>&1 | tail -8
echo "# tools/stereo_timing.py (C4)"; python tools/stereo_timing.py
echo "# tools/corner_timing.py"; python tools/corner_timing.py 64
echo "# tools/phase_clocks.py (VG_VARIANT=phase)"; VG_VARIANT=phase python tools/phase_clocks.py 10000 full
} > $O/r02_kernel_timing.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_bench_launch_list.csv python bench.py --steps 20 --warmup 3 --cpu-seconds 0.5 > $O/r02_bench_under_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r02_lm_solve_launch_list.csv python tools/lm_timing.py 10000 0 > /dev/null 2>&1
for cfg in "10k 0 10000" "25k 0 25000" "mei 2 10000"; do set -- $cfg
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:reproj -s 12 -c 1 -o $O/r02_eval_$1 -f python tools/kernel_timing.py --modes full --steps 20 --n-img $3 --model $2 > $O/r02_eval_$1.log 2>&1
  ncu -i $O/r02_eval_$1.ncu-rep --page raw --csv > $O/r02_eval_$1_raw.csv 2>/dev/null
done
timeout 600 ncu --set full --clock-control none -k regex:corner_response -s 3 -c 1 -o $O/r02_corner -f python tools/corner_timing.py 64 > $O/r02_corner.log 2>&1
ncu -i $O/r02_corner.ncu-rep --page raw --csv > $O/r02_corner_raw.csv 2>/dev/null
ls -la $O/r02_* | head -30

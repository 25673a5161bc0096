#!/bin/bash
# round-2 closing run on one GPU: the whole -m gpu suite, smoke(), the bench line with the driver's arguments, the
# reference arm, the launch list of the bench command, kernel / LM / detector timings
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/fin_tests.log 2>&1
tail -5 gpurun_out/fin_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fin_smoke.log 2>&1; tail -2 gpurun_out/fin_smoke.log
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/fin_bench_ref.json 2> gpurun_out/fin_bench_ref.err
python bench.py --steps 20 --warmup 5 > gpurun_out/fin_bench.json 2> gpurun_out/fin_bench.err
tail -c 1500 gpurun_out/fin_bench.json; tail -3 gpurun_out/fin_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/fin_launches.csv \
    python bench.py --steps 2 --warmup 1 --cpu-seconds 0.5 > gpurun_out/fin_bench_under_ncu.log 2>&1
( echo "# tools/kernel_timing.py"; python tools/kernel_timing.py --n-img 10000 --steps 300 --modes full,normal;
  python tools/kernel_timing.py --n-img 25000 --steps 200 --modes full,normal;
  python tools/kernel_timing.py --model 2 --n-img 10000 --steps 300 --modes full; python tools/kernel_timing.py --model 1 --n-img 10000 --steps 300 --modes full;
  echo "# VG_LATE_WAIT=0 (every launch waits at its head for the launch ahead)"; VG_LATE_WAIT=0 python tools/kernel_timing.py --n-img 10000 --steps 300 --modes full;
  echo "# tools/lm_timing.py"; python tools/lm_timing.py 2>&1 | tail -3;
  echo "# tools/stereo_timing.py"; python tools/stereo_timing.py 2>&1 | tail -3;
  echo "# tools/detector_timing.py 128"; python tools/detector_timing.py 128 2>&1 ) > gpurun_out/fin_timing.txt 2>&1
python tools/aux_timing.py > gpurun_out/fin_aux_timing.txt 2>&1
tail -30 gpurun_out/fin_timing.txt

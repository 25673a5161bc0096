set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_r1i.json 2> gpurun_out/bench_r1i.err; tail -c 400 gpurun_out/bench_r1i.json
python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_r1i_ref.json
python tools/kernel_timing.py 2>&1 | tail -4 > gpurun_out/timing_r1i.txt
python tools/kernel_timing.py --n-img 25000 --modes full,jac 2>&1 | tail -2 >> gpurun_out/timing_r1i.txt
python tools/kernel_timing.py --model 2 --modes full,normal 2>&1 | tail -2 >> gpurun_out/timing_r1i.txt
python tools/kernel_timing.py --model 1 --modes full,normal 2>&1 | tail -2 >> gpurun_out/timing_r1i.txt
python tools/lm_timing.py >> gpurun_out/timing_r1i.txt 2>&1
VG_LM_TRACE=1 python tools/lm_timing.py 2>&1 | tail -8 >> gpurun_out/timing_r1i.txt
cat gpurun_out/timing_r1i.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1i.csv python bench.py --steps 20 --warmup 3 --cpu-seconds 0.5 > gpurun_out/bench_under_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_lm_r1i.csv python tools/lm_timing.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:reproj -s 12 -c 1 -o gpurun_out/prof_r1i_10k python tools/kernel_timing.py --modes full --steps 20 --n-img 10000 > gpurun_out/ncu_r1i.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:reproj -s 12 -c 1 -o gpurun_out/prof_r1i_mei python tools/kernel_timing.py --model 2 --modes full --steps 20 --n-img 10000 > gpurun_out/ncu_r1i.log 2>&1

set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; tail -c 600 gpurun_out/bench_r1f.json
python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_r1f_ref.json
python tools/kernel_timing.py 2>&1 | tail -4 > gpurun_out/timing_r1f.txt
python tools/kernel_timing.py --n-img 25000 --modes full,jac 2>&1 | tail -2 >> gpurun_out/timing_r1f.txt
python tools/kernel_timing.py --model 2 --modes full,normal 2>&1 | tail -2 >> gpurun_out/timing_r1f.txt
python tools/kernel_timing.py --model 1 --modes full,normal 2>&1 | tail -2 >> gpurun_out/timing_r1f.txt
python tools/lm_timing.py >> gpurun_out/timing_r1f.txt 2>&1
cat gpurun_out/timing_r1f.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 20 --warmup 3 --cpu-seconds 0.5 > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:reproj -s 12 -c 1 -o gpurun_out/prof_r1f_10k python tools/kernel_timing.py --modes full --steps 20 --n-img 10000 > gpurun_out/ncu_r1f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:reproj -s 12 -c 1 -o gpurun_out/prof_r1f_25k python tools/kernel_timing.py --modes full --steps 20 --n-img 25000 > gpurun_out/ncu_r1f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:reproj -s 12 -c 1 -o gpurun_out/prof_r1f_mei python tools/kernel_timing.py --model 2 --modes full --steps 20 --n-img 10000 > gpurun_out/ncu_r1f.log 2>&1

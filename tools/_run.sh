set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/lm_timing.py 2>&1 | tail -2
python tools/lm_timing.py 10000 25 2>&1 | tail -1
python bench.py > gpurun_out/bench_r1g.json 2> gpurun_out/bench_r1g.err; tail -c 900 gpurun_out/bench_r1g.json; tail -3 gpurun_out/bench_r1g.err

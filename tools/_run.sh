set -x
python -m pytest tests -m gpu -q 2>&1 | tail -2
python bench.py > gpurun_out/bench_r1j.json 2> gpurun_out/bench_r1j.err; tail -c 300 gpurun_out/bench_r1j.json; tail -2 gpurun_out/bench_r1j.err

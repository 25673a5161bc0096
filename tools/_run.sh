timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 60 python tools/kernel_timing.py --modes full 2>&1 | tail -1
timeout 300 python bench.py 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_err.log; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}, d['roofline']['kernel_us'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['config']['cost_check'])
print(d['lm']); print(d['cpu_baseline'])
PY

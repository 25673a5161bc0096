#!/bin/bash
# last run of the round on one GPU: the whole -m gpu suite, smoke(), the reference arm and the bench line with the driver's arguments
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/fin_tests.log 2>&1
tail -5 gpurun_out/fin_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fin_smoke.log 2>&1; tail -1 gpurun_out/fin_smoke.log
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/fin_bench_ref.json 2> gpurun_out/fin_bench_ref.err
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/fin_bench.json 2> gpurun_out/fin_bench.err
tail -4 gpurun_out/fin_bench.err
python - <<PY
import json
b=json.loads([l for l in open('gpurun_out/fin_bench.json') if l.startswith('{')][-1])
r=json.loads([l for l in open('gpurun_out/fin_bench_ref.json') if l.startswith('{')][-1])
print('value %.3f G/s  step %.2f us  kernel %.2f us frac %.3f step_frac %.3f  lm %.0f  e2e %.2f G/s (params only %.2f) ceres %.0f / %.0f M/s  cpu %.1f M/s det %.0f  ref arm %.1f M/s' % (b['value']/1e9, b['ms_per_step']*1e3, b['roofline']['kernel_us'], b['roofline']['frac'], b['roofline']['step_frac'], b['lm']['iters_per_s'], b['e2e']['value']/1e9, b['e2e']['parameters_only']['value']/1e9, b['e2e_ceres_contract']['value']/1e6, b['e2e_ceres_contract']['value_with_registered_outputs']/1e6, b['cpu_baseline']['value']/1e6, b['detector']['images_per_s'], r['value']/1e6))
PY

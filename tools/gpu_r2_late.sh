#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_eval_gpu.py tests/test_solve_gpu.py tests/test_peer_exchange_gpu.py tests/test_golden_gpu.py tests/test_priors_gpu.py -x -q -m gpu > gpurun_out/late_tests.log 2>&1
tail -4 gpurun_out/late_tests.log
for v in 1 0; do
echo "== VG_LATE_WAIT=$v"
VG_LATE_WAIT=$v python bench.py --steps 20 --warmup 5 --cpu-seconds 0.5 > gpurun_out/late_bench_$v.json 2> gpurun_out/late_bench_$v.err
python - <<PY
import json
b=json.loads([l for l in open('gpurun_out/late_bench_$v.json') if l.startswith('{')][-1])
print('value %.3f G/s  step %.2f us  kernel %.2f us frac %.3f step_frac %.3f  lm %.0f  e2e %.2f G/s  cost %r' % (b['value']/1e9, b['ms_per_step']*1e3, b['roofline']['kernel_us'], b['roofline']['frac'], b['roofline']['step_frac'], b['lm']['iters_per_s'], b['e2e']['value']/1e9, b['details']['cost_check']))
PY
VG_LATE_WAIT=$v python bench.py --steps 200 --warmup 10 --cpu-seconds 0.5 2>/dev/null | python -c "
import json,sys
b=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('K=200: value %.3f G/s step %.2f us kernel %.2f us' % (b['value']/1e9, b['ms_per_step']*1e3, b['roofline']['kernel_us']))"
done

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_eval_gpu.py -x -q -m gpu -k "registered or chunking or concurrent" > gpurun_out/reg_tests.log 2>&1
tail -5 gpurun_out/reg_tests.log
python bench.py --steps 20 --warmup 5 --cpu-seconds 1 > gpurun_out/reg_bench.json 2> gpurun_out/reg_bench.err
python - <<'PY'
import json
b=json.loads([l for l in open('gpurun_out/reg_bench.json') if l.startswith('{')][-1])
print(b['e2e_ceres_contract']['value'], b['e2e_ceres_contract']['value_with_registered_outputs'], b['e2e']['value'], b['value'])
PY
tail -2 gpurun_out/reg_bench.err

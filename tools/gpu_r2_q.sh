#!/bin/bash
# round 2, GPU call Q (2 GPUs): peer-exchange tests, bench.py at N=2 with the exchange check
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_peer_exchange_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r2q_bench2.json 2> gpurun_out/r2q_bench2.err; echo "bench rc $?"; tail -c 600 gpurun_out/r2q_bench2.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/r2q_bench2.json").read().strip().splitlines()[-1])
print({k:l.get(k) for k in ("value","ms_per_step","exchange_check","failures")}); print(l["lm"]); print(l["e2e"])
PY

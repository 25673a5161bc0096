#!/bin/bash
# 2 GPUs: deferred exchange posted by the next launch's head (default for one process per rank) vs by the tail
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_peer_exchange_gpu.py -m gpu -x -q 2>&1 | tail -3
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 --cpu-seconds 0.5 2>/dev/null | tail -1 > $1; python - <<PY
import json
l=json.loads(open("$1").read().strip().splitlines()[-1])
print("$1", l["ms_per_step"], l["value"], l["roofline"]["kernel_us"], l["lm"]["iters_per_s"], l["exchange_check"]["ok"], l["exchange_check"]["max_rel_vs_one_gpu_on_all_images"])
PY
}
run gpurun_out/r2x_head_1.json
VG_PEER_POST_TAIL=1 run gpurun_out/r2x_tail_1.json
run gpurun_out/r2x_head_2.json
VG_PEER_POST_TAIL=1 run gpurun_out/r2x_tail_2.json
timeout 300 python bench.py --steps 100 --warmup 10 --cpu-seconds 0.5 2>/dev/null | tail -1 > gpurun_out/r2x_one.json
python - <<PY
import json
l=json.loads(open("gpurun_out/r2x_one.json").read().strip().splitlines()[-1])
print("one gpu", l["ms_per_step"], l["value"], l["roofline"]["kernel_us"], l["lm"]["iters_per_s"])
PY

#!/bin/bash
# round 2, GPU call AC (8 GPUs): bench.py --gpus 8 with the exchange check and the C5 leg
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/r2ac_bench8.json 2> gpurun_out/r2ac_bench8.err; echo "bench rc $?"; tail -c 800 gpurun_out/r2ac_bench8.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/r2ac_bench8.json").read().strip().splitlines()[-1])
print({k:l.get(k) for k in ("value","ms_per_step","failures")}); print("check", l["exchange_check"]["ok"], l["exchange_check"]["max_rel"], l["exchange_check"]["max_rel_vs_one_gpu_on_all_images"]); print("lm", l["lm"]["iters_per_s"], l["lm"]["iterations"], l["lm"]["bit_identical_across_ranks"]); print("e2e", l["e2e"]["value"])
c=l.get("c5"); print("c5", None if c is None else {k:c[k] for k in ("value","ms_per_step","images_total")}, None if c is None else c["roofline"]["frac"], None if c is None else c["roofline"]["step_frac"], None if c is None else c["lm"], None if c is None else c["exchange_check"]["ok"])
PY

set -x
timeout 300 python -m pytest tests/test_peer_exchange_gpu.py -x -q --tb=short 2>&1 | grep -B2 -A25 "WORKER FAILED\|passed" | head -40
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 2>gpurun_out/bench2.err | tail -1 > gpurun_out/bench_r1i_2gpu.json; python -c "
import json; d=json.loads(open('gpurun_out/bench_r1i_2gpu.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['lm']['iters_per_s'], d['lm']['iterations'], d['lm']['final_cost'], d['config']['cost_check'])"; tail -2 gpurun_out/bench2.err

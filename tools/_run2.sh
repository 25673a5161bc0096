set -x
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 100 --warmup 10 2>gpurun_out/bench4.err | tail -1 > gpurun_out/bench_r1i_4gpu.json; python -c "
import json; d=json.loads(open('gpurun_out/bench_r1i_4gpu.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['lm']['iters_per_s'], d['lm']['iterations'], d['lm']['final_cost'], d['config']['collective'])"; tail -2 gpurun_out/bench4.err

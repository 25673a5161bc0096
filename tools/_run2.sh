set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 2>gpurun_out/bench2.err | tail -1 > gpurun_out/bench_r1g_2gpu.json; tail -c 1800 gpurun_out/bench_r1g_2gpu.json; tail -3 gpurun_out/bench2.err

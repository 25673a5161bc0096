#!/bin/bash
# round-2 closing run on N GPUs (N = first argument): the peer-exchange tests, then the bench line
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_peer_exchange_gpu.py -x -q -m gpu > gpurun_out/fin${N}_tests.log 2>&1
tail -3 gpurun_out/fin${N}_tests.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/fin${N}_bench.json 2> gpurun_out/fin${N}_bench.err
tail -c 3000 gpurun_out/fin${N}_bench.json; tail -3 gpurun_out/fin${N}_bench.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/fin${N}_bench_ref.json 2> gpurun_out/fin${N}_bench_ref.err
tail -c 400 gpurun_out/fin${N}_bench_ref.json

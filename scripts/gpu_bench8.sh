#!/bin/bash
# bench.py at N GPUs only (no tests): $1 = N, $2 = tag
N=${1:-8}
TAG=${2:-r01}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$TAG.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 \
    > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err
cat gpurun_out/bench_n${N}_$TAG.json; tail -5 gpurun_out/bench_n${N}_$TAG.err

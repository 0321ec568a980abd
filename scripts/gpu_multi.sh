#!/bin/bash
# multi-GPU round: all GPU parity tests (incl. whole-genome mode at 2 ranks), bench.py at 1 and N GPUs ($1, default 2)
N=${1:-2}
TAG=${2:-r01}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$TAG.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/test_multi_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_multi_$TAG.log
tail -15 gpurun_out/test_multi_$TAG.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err
cat gpurun_out/bench_n1_$TAG.json; tail -3 gpurun_out/bench_n1_$TAG.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 \
    > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err
cat gpurun_out/bench_n${N}_$TAG.json; tail -5 gpurun_out/bench_n${N}_$TAG.err

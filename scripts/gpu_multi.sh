#!/bin/bash
# multi-GPU round: whole-genome-mode parity tests at N ranks, bench.py at N GPUs ($1, default 2)
N=${1:-2}
TAG=${2:-r01}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$TAG.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/test_multi_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_multi_$TAG.log
tail -6 gpurun_out/test_multi_$TAG.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 \
    > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err
cat gpurun_out/bench_n${N}_$TAG.json; tail -5 gpurun_out/bench_n${N}_$TAG.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 \
    > gpurun_out/bench_ref_n${N}_$TAG.json 2> gpurun_out/bench_ref_n${N}_$TAG.err
cat gpurun_out/bench_ref_n${N}_$TAG.json | cut -c1-600; tail -3 gpurun_out/bench_ref_n${N}_$TAG.err

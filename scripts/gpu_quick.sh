#!/bin/bash
# quick GPU check: parity tests + bench (no ncu). $1 = tag
mkdir -p gpurun_out
TAG=${1:-q}
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/test_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_$TAG.log
tail -4 gpurun_out/test_$TAG.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err

#!/bin/bash
# compute-sanitizer on the device decode tests. $1 = tag
TAG=${1:-b26}
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_bamdev.py tests/test_zz_gpu_inflate.py -m gpu -x -q > gpurun_out/memcheck_$TAG.log 2>&1
tail -4 gpurun_out/memcheck_$TAG.log
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_bamdev.py tests/test_zz_gpu_inflate.py -m gpu -x -q -k "level and 6 or decoy or overlapping" > gpurun_out/racecheck_$TAG.log 2>&1
tail -4 gpurun_out/racecheck_$TAG.log

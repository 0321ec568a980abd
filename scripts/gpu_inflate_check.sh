#!/bin/bash
# GPU check of the opt-in device inflate: kernel tests, CLI goldens (host decoder changes), decode timing with and without it.
# $1 = tag
mkdir -p gpurun_out
TAG=${1:-r03a}
timeout 120 python -m pytest tests/test_zz_gpu_inflate.py -m gpu -x -q > gpurun_out/test_inflate_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_inflate_$TAG.log
tail -15 gpurun_out/test_inflate_$TAG.log
timeout 90 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k cli > gpurun_out/test_cli_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_cli_$TAG.log
tail -3 gpurun_out/test_cli_$TAG.log
nproc > gpurun_out/decode_$TAG.txt
PAIRS=${2:-1000000}
timeout 120 env BDK_DECODE_TRACE=1 python scripts/bam_decode_bench.py --pairs $PAIRS --reps 3 >> gpurun_out/decode_$TAG.txt 2>&1
timeout 60 env BDK_DECODE_TRACE=1 BDK_GPU_INFLATE=1 python scripts/bam_decode_bench.py --pairs $PAIRS --reps 3 >> gpurun_out/decode_$TAG.txt 2>&1
grep -v "record chain:" gpurun_out/decode_$TAG.txt | tail -12

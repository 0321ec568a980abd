#!/bin/bash
# device-resident BAM decode: its GPU tests, the inflate tests, the CLI tests, then the measurement. $1 = tag, $2 = pairs
mkdir -p gpurun_out
TAG=${1:-b01}
timeout 900 python -m pytest tests/test_gpu_bamdev.py tests/test_zz_gpu_inflate.py -m gpu -x -q -s --durations=8 > gpurun_out/test_bamdev_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_bamdev_$TAG.log
tail -30 gpurun_out/test_bamdev_$TAG.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cli" > gpurun_out/test_cli_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_cli_$TAG.log
tail -5 gpurun_out/test_cli_$TAG.log
timeout 900 python scripts/bamdev_bench.py ${2:-6000000} 6 > gpurun_out/bamdev_$TAG.json 2> gpurun_out/bamdev_$TAG.err
cat gpurun_out/bamdev_$TAG.json; tail -5 gpurun_out/bamdev_$TAG.err
if [ -n "$NCU" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bgzf_inflate_warp -s 3 -c 1 -o gpurun_out/inflate_warp_$TAG -f python scripts/bamdev_bench.py 2000000 6 > gpurun_out/ncu_inflate_$TAG.log 2>&1
ls -la gpurun_out | grep inflate_warp
fi

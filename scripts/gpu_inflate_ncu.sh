#!/bin/bash
# ncu --set full of one inflate launch that fills the GPU (one chunk = one window of 512 MiB of output: ~8000 members on 592 CTAs). $1 = tag
mkdir -p gpurun_out
TAG=${1:-b13}
export BDK_BAMDEV_WINDOW_KB=262144 BDK_BAMDEV_CHUNK_KB=262144
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bgzf_inflate_warp -s 1 -c 1 -o gpurun_out/inflate_warp_$TAG -f python scripts/bamdev_bench.py ${2:-6000000} 6 > gpurun_out/ncu_inflate_$TAG.log 2>&1
tail -3 gpurun_out/ncu_inflate_$TAG.log
ls -la gpurun_out | grep inflate_warp_$TAG

#!/bin/bash
# GPU round: parity tests, bench, ncu launch list, ncu full capture of one kernel ($1 = kernel regex, default k1_classify)
mkdir -p gpurun_out
TAG=${2:-r01}
K=${1:-k1_classify}
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/test_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_$TAG.log
tail -5 gpurun_out/test_$TAG.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -o gpurun_out/${K}_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out

#!/bin/bash
# GPU round: parity tests, bench, ncu launch list, ncu full capture of one kernel ($1 = kernel regex, default k1_classify), $2 = tag
# env: SKIP_NCU=1 no profiler passes; CONFIG3=1 also the configs[2] bench + K4 traces; SKIP_TESTS=1
mkdir -p gpurun_out
TAG=${2:-r01}
K=${1:-k1_classify}
if [ -z "$SKIP_TESTS" ]; then
timeout 1500 python -m pytest tests -m gpu -x -q -s --durations=12 > gpurun_out/test_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_$TAG.log
grep -E "config 3 full size|passed|failed|rc=|Error|error|s call|s setup" gpurun_out/test_$TAG.log | tail -22
fi
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
if [ -n "$CONFIG3" ]; then
timeout 900 python bench.py --config 3 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err
cat gpurun_out/bench_c3_$TAG.json; tail -3 gpurun_out/bench_c3_$TAG.err
timeout 600 python scripts/k4_trace.py 3 > gpurun_out/k4trace_c3_$TAG.log 2>&1; tail -25 gpurun_out/k4trace_c3_$TAG.log
timeout 600 python scripts/k4_trace.py 2 > gpurun_out/k4trace_c2_$TAG.log 2>&1; tail -14 gpurun_out/k4trace_c2_$TAG.log
fi
if [ -z "$SKIP_NCU" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -o gpurun_out/${K}_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full_$TAG.log 2>&1
fi
ls -la gpurun_out | tail -12
if [ -n "$K1C3" ]; then   # full capture of the multi-key classify kernel on configs[2] data
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_classify -c 1 -o gpurun_out/k1_config3_$TAG -f python bench.py --config 3 --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_k1c3_$TAG.log 2>&1
ls -la gpurun_out | grep k1_config3
fi

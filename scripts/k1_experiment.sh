#!/bin/bash
# K1 experiment: bench both configs without the CPU legs
mkdir -p gpurun_out
TAG=$1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
print("config2", d["ms_per_step"], d["roofline"]["frac"], d["kernel_ms_per_step"]["k1_classify"], d["e2e"]["ms_per_step"])
PY
timeout 900 python bench.py --config 3 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_c3_$TAG.json").read().strip().splitlines()[-1])
print("config3", d["ms_per_step"], d["roofline"]["frac"], d["kernel_ms_per_step"])
PY
tail -3 gpurun_out/bench_$TAG.err gpurun_out/bench_c3_$TAG.err

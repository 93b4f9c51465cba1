#!/bin/bash
# tests + bench (+ per-kernel profile) + pipeline trace + config 3/4/5 throughput, one GPU
set -u
mkdir -p gpurun_out
TAG=${1:-q}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_$TAG.log
python bench.py --no-cpu-baseline --steps 30 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json")); print("bench", d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"]); print(d["roofline"]["per_kernel_ms_per_step"])
PY
tail -3 gpurun_out/bench_$TAG.err
python tools/trace_run.py "23,43,44,47" gpurun_out/tc_trace_$TAG.txt 2>&1 | tail -1
python tools/bench_configs.py --configs 3,4,5 --steps 3 2>&1 | tail -3

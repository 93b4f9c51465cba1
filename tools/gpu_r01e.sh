#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r01e}
./tools/launch_bench > gpurun_out/launch_bench_$TAG.log 2>&1; cat gpurun_out/launch_bench_$TAG.log
./tools/sync_bench > gpurun_out/sync_bench_$TAG.log 2>&1; cat gpurun_out/sync_bench_$TAG.log
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json")); print("bench", d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"]); print(d["roofline"]["per_kernel_ms_per_step"])
PY
tail -3 gpurun_out/bench_$TAG.err

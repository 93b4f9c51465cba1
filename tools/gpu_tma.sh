#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --no-cpu-baseline --steps 30 > gpurun_out/bench_tma1.json 2>gpurun_out/bench_ab.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_tma1.json")); k=d["roofline"]["per_kernel_ms_per_step"]
print("bench", d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"]); print(k)
PY
tail -2 gpurun_out/bench_ab.err
python tools/trace_run.py "23,43,44,47" gpurun_out/tc_trace_tma.txt 2>&1 | tail -1
python tools/bench_configs.py --configs 3,4,5 --steps 3 2>&1 | tail -3

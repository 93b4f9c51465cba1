#!/bin/bash
mkdir -p gpurun_out
for o in "weight_prefetch=1" "weight_prefetch=0"; do
  TVC_OPTS=$o python bench.py --no-cpu-baseline --steps 30 > gpurun_out/bench_ab.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_ab.json")); print("$o", d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["roofline"]["per_kernel_ms_per_step"].get("weights_to_l2"))
PY
done
python -m pytest tests/test_gpu_decoder.py -m gpu -x -q 2>&1 | tail -2

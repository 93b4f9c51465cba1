#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encoder_knn.py tests/test_gpu_configs.py -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_r02p.log
timeout 600 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/bench_r02p.json 2> gpurun_out/bench_r02p.err; tail -3 gpurun_out/bench_r02p.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r02p.json"))
print("config2", d["ms_per_step"], d["value"]/1e6, d["roofline"]["frac"], "e2e", d["e2e"]["value"]/1e6)
for k,v in d["other_configs"].items(): print(k, {a:b for a,b in v.items() if a!="workload"})
PY

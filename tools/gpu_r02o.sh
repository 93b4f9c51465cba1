#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encoder_knn.py -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_r02o.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/parity_report.json"))
for k in ("frontend","encoder","pipeline_stages","config3"):
    if k in d: print(k, d[k])
PY
python tools/profile_tick.py 128 2>&1 | tee gpurun_out/r02o_profile_tick.log
timeout 600 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/bench_r02o.json 2> gpurun_out/bench_r02o.err; tail -3 gpurun_out/bench_r02o.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r02o.json"))
print("config2", d["ms_per_step"], d["value"]/1e6, d["roofline"]["frac"], "e2e", d["e2e"]["value"]/1e6)
for k,v in d["other_configs"].items(): print(k, {a:b for a,b in v.items() if a!="workload"})
PY

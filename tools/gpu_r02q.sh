#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python - <<'PY'
import json
d=json.load(open("gpurun_out/parity_report.json"))
for k,v in d.items():
    if "decoder" in k or "rmse" in str(v): print(k, v)
PY
for cat in 0 1 0 1; do
  TVC_TC_CAT=$cat timeout 300 python bench.py --no-cpu-baseline --no-extra-configs --steps 40 > gpurun_out/bench_ab_tmp.json 2>gpurun_out/bench_ab_tmp.err
  python - $cat <<'PY'
import json, sys
d = json.load(open("gpurun_out/bench_ab_tmp.json")); k = d["roofline"]["breakdown"]["per_kernel_ms_per_step"]
top = sorted(k.items(), key=lambda kv: -kv[1])[:6]
print(f"cat={sys.argv[1]} ms_per_step={d['ms_per_step']:.4f} value={d['value']/1e6:.1f}M", {n: round(v, 4) for n, v in top})
PY
done
for cat in 0 1; do TVC_TC_CAT=$cat python tools/profile_shape.py 64 500 2>&1 | head -4; done

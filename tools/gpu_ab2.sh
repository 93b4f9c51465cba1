#!/bin/bash
# same-box A/B: two MMA issuers everywhere (1) vs on single-stage tiles only (2) vs off (0)
mkdir -p gpurun_out
for m in 1 2 0 2 1; do
  TVC_TC_MMA2=$m python bench.py --no-cpu-baseline --steps 40 > gpurun_out/bench_ab.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_ab.json")); k=d["roofline"]["per_kernel_ms_per_step"]
print("mma2=$m", round(d["ms_per_step"],4), round(d["value"]/1e6,1), {n:k[n] for n in ("tc_up4_c1","tc_up4_c2","tc_up4_c4","tc_up4_c5","tc_up3_c2","tc_down0")})
PY
done 2>&1 | tee gpurun_out/ab2.log

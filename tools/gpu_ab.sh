#!/bin/bash
# same-box A/B of the tc_conv switches (box-to-box variance is +-3 %, larger than most of these effects)
mkdir -p gpurun_out
for cfg in "1 1" "0 1" "1 0" "0 0" "1 1" "0 0"; do
  set -- $cfg
  TVC_TC_MMA2=$1 TVC_TC_WRES=$2 python bench.py --no-cpu-baseline --steps 40 > gpurun_out/bench_ab.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_ab.json")); k=d["roofline"]["per_kernel_ms_per_step"]
print("mma2=$1 wres=$2", round(d["ms_per_step"],4), round(d["value"]/1e6,1), {n:k[n] for n in ("tc_up4_c1","tc_up4_c2","tc_up4_c5","tc_up3_c1","tc_down0")})
PY
done
TVC_TC_MMA2=0 TVC_TC_WRES=0 python tools/bench_configs.py --configs 4 --steps 3 2>&1 | tail -1 | cut -c1-200
TVC_TC_MMA2=1 TVC_TC_WRES=1 python tools/bench_configs.py --configs 4 --steps 3 2>&1 | tail -1 | cut -c1-200

#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for t in 1 0; do
  TVC_TC_TMA=$t python bench.py --no-cpu-baseline --steps 30 > gpurun_out/bench_tma$t.json 2>gpurun_out/bench_ab.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_tma$t.json")); k=d["roofline"]["per_kernel_ms_per_step"]
print("tma=$t", d["ms_per_step"], d["value"], {n:k[n] for n in ("tc_up4_c1","tc_up4_c2","tc_up4_c5","tc_up3_c1","tc_up2_c1","tc_up0_c1","tc_up0_c5","tc_down0","tc_idft","tc_heads","tc_cnxt_c2","tc_frame_in")})
PY
  tail -2 gpurun_out/bench_ab.err
done
python tools/bench_configs.py --configs 3,4 --steps 3 2>&1 | tail -2

#!/bin/bash
mkdir -p gpurun_out
TVC_TC_BULK=1 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_tc_conv.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -3
for b in 1 0; do
  TVC_TC_BULK=$b python bench.py --no-cpu-baseline --steps 30 > gpurun_out/bench_ab.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_ab.json")); k=d["roofline"]["per_kernel_ms_per_step"]
print("bulk=$b", d["ms_per_step"], d["value"], {n:k[n] for n in ("tc_up4_c1","tc_up4_c2","tc_up4_c5","tc_up3_c1","tc_up0_c1","tc_up0_c5","tc_down0","tc_idft","tc_heads","tc_cnxt_c2")})
PY
done
TVC_TC_BULK=1 python tools/trace_run.py "43,47" gpurun_out/tc_trace_bulk.txt 2>&1 | tail -1
TVC_TC_BULK=1 python tools/bench_configs.py --configs 4 --steps 3 2>&1 | tail -1

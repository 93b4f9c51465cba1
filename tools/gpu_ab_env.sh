#!/bin/bash
# Same-box A/B of environment settings: tools/gpu_ab_env.sh "VAR=1" "VAR=2" ...   ("" = defaults); alternates twice
mkdir -p gpurun_out
for rep in 1 2; do
for e in "$@"; do
  env $e timeout 300 python bench.py --no-cpu-baseline --no-extra-configs --steps 40 > gpurun_out/bench_ab_tmp.json 2>gpurun_out/bench_ab_tmp.err
  python - "$e" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/bench_ab_tmp.json")); k = d["roofline"]["breakdown"]["per_kernel_ms_per_step"]
    grp = lambda p: sum(v for n, v in k.items() if n.startswith(p))
    print(f"env={sys.argv[1]!r:24s} ms_per_step={d['ms_per_step']:.4f}", {g: round(grp(g), 4) for g in ("tc_up0", "tc_up1", "tc_up2", "tc_up3", "tc_down", "tc_cnxt", "tc_idft", "tc_frame_in", "tc_heads")})
except Exception as ex:
    print(f"env={sys.argv[1]!r} FAILED {ex}"); print(open("gpurun_out/bench_ab_tmp.err").read()[-1500:])
PY
done
done

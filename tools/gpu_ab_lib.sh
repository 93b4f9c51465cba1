#!/bin/bash
# Same-box A/B of library builds: tools/gpu_ab_lib.sh libA.so libB.so ...   ("" = the in-tree library); alternates twice
mkdir -p gpurun_out
for rep in 1 2; do
for lib in "$@"; do
  TVC_LIB="$lib" timeout 300 python bench.py --no-cpu-baseline --no-extra-configs --steps 40 > gpurun_out/bench_ab_tmp.json 2>gpurun_out/bench_ab_tmp.err
  python - "$lib" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/bench_ab_tmp.json")); k = d["roofline"]["breakdown"]["per_kernel_ms_per_step"]
    top = sorted(k.items(), key=lambda kv: -kv[1])[:5]
    print(f"lib={sys.argv[1] or 'in-tree'!r:40s} ms_per_step={d['ms_per_step']:.4f} value={d['value']/1e6:.1f}M e2e={d['e2e']['value']/1e6:.1f}M", {n: round(v, 4) for n, v in top})
except Exception as e:
    print(f"lib={sys.argv[1]!r} FAILED {e}"); print(open("gpurun_out/bench_ab_tmp.err").read()[-1500:])
PY
done
done

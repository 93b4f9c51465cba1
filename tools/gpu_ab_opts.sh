#!/bin/bash
# Same-box A/B of plan options: tools/gpu_ab_opts.sh "chain=0" "chain=1" ...   (each arg = a TVC_OPTS string; "" = defaults)
mkdir -p gpurun_out
for opts in "$@"; do
  TVC_OPTS="$opts" timeout 300 python bench.py --no-cpu-baseline --no-extra-configs --steps 40 > gpurun_out/bench_ab_tmp.json 2>gpurun_out/bench_ab_tmp.err
  python - "$opts" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/bench_ab_tmp.json")); k = d["roofline"]["breakdown"]["per_kernel_ms_per_step"]
    top = sorted(k.items(), key=lambda kv: -kv[1])[:8]
    print(f"opts={sys.argv[1]!r} ms_per_step={d['ms_per_step']:.4f} value={d['value']/1e6:.1f}M e2e={d['e2e']['value']/1e6:.1f}M launches/step={d['gpu_launches']/d['steps']:.0f}", {n: round(v, 4) for n, v in top})
except Exception as e:
    print(f"opts={sys.argv[1]!r} FAILED {e}"); print(open("gpurun_out/bench_ab_tmp.err").read()[-2000:])
PY
done

#!/bin/bash
# per-layer ablation: TVC_TC_DBG bit 1 = no operand loads, 2 = no MMAs, 4 = no epilogue math / stores
mkdir -p gpurun_out
for d in 0 1 2 3 4 7; do
  TVC_TC_DBG=$d TVC_OPTS="pad_max_t=0" timeout 300 python bench.py --no-cpu-baseline --no-extra-configs --steps 20 > gpurun_out/abl_$d.json 2>gpurun_out/abl_$d.err || tail -3 gpurun_out/abl_$d.err
done
python - <<'PY'
import json
ds=[0,1,2,3,4,7]
t={d:json.load(open(f"gpurun_out/abl_{d}.json")) for d in ds}
print("step ms:", {d: round(t[d]["ms_per_step"],4) for d in ds})
k0=t[0]["roofline"]["breakdown"]["per_kernel_ms_per_step"]
print(f"{'kernel':18s}"+"".join(f"{'dbg'+str(d):>8s}" for d in ds))
for k in sorted(k0, key=lambda k:-k0[k]):
    print(f"{k:18s}"+"".join(f"{t[d]['roofline']['breakdown']['per_kernel_ms_per_step'].get(k,0)*1e3:8.1f}" for d in ds))
PY

#!/usr/bin/env python
"""Turn the files a `tools/gpu_final.sh <tag>` run left in gpurun_out/ into the small, tracked artefacts under profiles/:
one step of the ncu launch list, roofline_traffic.json (DRAM bytes of the conv launches of that step), summaries of the
full ncu captures, and copies of the bench / parity / pytest outputs."""
import csv
import json
import shutil
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02z"
rows = list(csv.reader(open(f"gpurun_out/launches_{tag}.csv")))
i0 = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[i0]
names = {}
for r in rows[i0 + 1:]:
    if len(r) == len(h):
        names[int(r[0])] = r[h.index("Kernel Name")]
fp = [i for i, n in sorted(names.items()) if "frame_prep" in n]
a, b = fp[3], fp[4]                      # one Decoder.infer step = from one frame_prep launch to the next
out = [h] + [r for r in rows[i0 + 1:] if len(r) == len(h) and a <= int(r[0]) < b]
with open(f"profiles/{tag}_launches_one_step.csv", "w", newline="") as f:
    csv.writer(f).writerows(out)
per = {}
for r in out[1:]:
    e = per.setdefault(int(r[0]), {"name": r[h.index("Kernel Name")]})
    v, u, m = float(r[h.index("Metric Value")]), r[h.index("Metric Unit")], r[h.index("Metric Name")]
    if "bytes" in m:
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    if "time" in m:
        v *= {"ns": 1e-3, "us": 1, "ms": 1e3}[u]
    e[m] = v
own = {k: e for k, e in per.items() if "tvc" in e["name"] or "unnamed" in e["name"]}
conv = [e for e in own.values() if "tc_conv_kernel" in e["name"] or "tc_up24_block" in e["name"]]
dram = lambda e: e["dram__bytes_read.sum"] + e["dram__bytes_write.sum"]
t_conv = sum(e["gpu__time_duration.sum"] for e in conv)
t_all = sum(e["gpu__time_duration.sum"] for e in own.values())
json.dump({"source": f"profiles/{tag}_launches_one_step.csv (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
                     "dram__bytes_write.sum --clock-control none on `bench.py --steps 2 --warmup 3`; one Decoder.infer step; "
                     "cold-cache, serialised)",
           "decoder_dram_bytes_per_step": sum(map(dram, own.values())),
           "dense_conv_dram_bytes_per_step": sum(map(dram, conv)),
           "dense_conv_launches_per_step": len(conv), "launches_per_step": len(own), "dense_conv_time_us_serialised": t_conv,
           "step_time_us_serialised": t_all, "dense_conv_share_serialised": t_conv / t_all},
          open("profiles/roofline_traffic.json", "w"), indent=1)
print(len(own), "launches,", len(conv), "dense convs, share", round(t_conv / t_all, 3))
import os
for t in ("block", "up3c2", "encgemm", "knn", "fft"):
    rep = f"gpurun_out/prof_{tag}_{t}.ncu-rep"
    if os.path.exists(rep):
        subprocess.run([sys.executable, "tools/ncu_summary.py", rep, f"profiles/{tag}_{t}_ncu_full.csv"], stdout=subprocess.DEVNULL)
for src, dst in ((f"bench_{tag}.json", f"{tag}_bench.json"), (f"bench_ref_{tag}.json", f"{tag}_bench_reference_cpu.json"),
                 ("parity_report.json", f"{tag}_parity_report.json"), (f"pytest_gpu_{tag}.log", f"{tag}_pytest_gpu.log"),
                 (f"tick_latency_{tag}.log", f"{tag}_tick_latency.log"), (f"profile_tick_{tag}.log", f"{tag}_profile_tick.log"),
                 (f"profile_c4_{tag}.log", f"{tag}_profile_c4.log"), (f"tick_launches_{tag}.csv", f"{tag}_tick_launches.csv")):
    if os.path.exists(f"gpurun_out/{src}"):
        shutil.copy(f"gpurun_out/{src}", f"profiles/{dst}")

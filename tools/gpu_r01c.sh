#!/bin/bash
# tests + bench (pdl on/off) + full ncu captures of the full-rate kernels
set -u
mkdir -p gpurun_out
TAG=${1:-r01c}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json")); print("pdl=1", d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"])
PY
TVC_OPTS=pdl=0 python bench.py --no-cpu-baseline > gpurun_out/bench_${TAG}_nopdl.json 2>> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}_nopdl.json")); print("pdl=0", d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"])
PY
tail -3 gpurun_out/bench_$TAG.err
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:tc_conv_kernel|osc_source|interp_cl|out_conv_k7' -s 52 -c 7 \
    -f -o gpurun_out/prof_${TAG}_up4 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu up4 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:osc_source|osc_frame_sums|tc_conv_kernel<4>' -c 3 \
    -f -o gpurun_out/prof_${TAG}_src python bench.py --steps 1 --warmup 3 --no-cpu-baseline >> gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu src rc=$?"
ls -la gpurun_out | tail -12

#!/bin/bash
# Round artefacts: GPU parity tests, both bench arms, ncu launch list of one bench command, full captures of the
# dominant decoder kernels and of the kNN screening GEMM.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r01z}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; echo "ref rc=$?"
cut -c1-300 gpurun_out/bench_ref_$TAG.json
python tools/bench_configs.py --configs 3,4,5 --steps 3 > gpurun_out/configs_$TAG.jsonl 2>> gpurun_out/bench_$TAG.err; cat gpurun_out/configs_$TAG.jsonl
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list_$TAG.log 2>&1
echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 43 -c 5 \
    -f -o gpurun_out/prof_${TAG}_up4 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu up4 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -c 1 \
    -f -o gpurun_out/prof_${TAG}_knn python tools/bench_configs.py --configs 3 --steps 1 >> gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu knn rc=$?"
ls -la gpurun_out | tail -15

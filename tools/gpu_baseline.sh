#!/bin/bash
# One gpurun call: GPU parity tests, contract bench, ncu launch list and one full capture of the
# dense-conv kernels.  Outputs (small) go to gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r1}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_list_$TAG.log 2>&1
echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:${NCU_K:-conv1d} -s ${NCU_S:-120} -c ${NCU_C:-4} \
    -f -o gpurun_out/prof_$TAG python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out

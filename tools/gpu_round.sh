#!/bin/bash
# One gpurun call: GPU parity tests, contract bench (both arms), ncu launch list of one bench command and one
# full capture of the dominant kernel.  Outputs (small) go to gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r1}
if [ "${NO_TESTS:-0}" != "1" ]; then
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_$TAG.log
fi
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
if [ "${NO_REF:-0}" != "1" ]; then
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; echo "ref rc=$?"
cat gpurun_out/bench_ref_$TAG.json
fi
if [ "${NO_LIST:-0}" != "1" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c ${LIST_C:-1500} --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list_$TAG.log 2>&1
echo "ncu list rc=$?"; tail -2 gpurun_out/ncu_list_$TAG.log
fi
if [ -n "${NCU_K:-}" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$NCU_K -s ${NCU_S:-0} -c ${NCU_C:-4} \
    -f -o gpurun_out/prof_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full_$TAG.log
fi
ls -la gpurun_out | head -30

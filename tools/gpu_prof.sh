#!/bin/bash
# bench + ncu launch list + one full capture (kernel regex / skip / count from env) -> gpurun_out/
set -u
TAG=${1:-x}
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json
if [ "${NO_LIST:-0}" != "1" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s ${LIST_S:-260} -c ${LIST_C:-90} --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list_$TAG.log 2>&1
echo "ncu list rc=$?"
fi
if [ -n "${NCU_K:-}" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$NCU_K -s ${NCU_S:-0} -c ${NCU_C:-3} \
    -f -o gpurun_out/prof_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full_$TAG.log
fi
ls -la gpurun_out | head -30

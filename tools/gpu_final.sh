#!/bin/bash
# Round artefacts: GPU parity tests, both bench arms, the ncu launch list of one bench command, full ncu captures of the
# dominant kernels (fused full-rate block, a mid-rate FiLM conv, the Encoder's GEMM, the kNN screening GEMM with its fused
# top-k, the FFT).  Everything lands in gpurun_out/; tools/collect_profiles.py <tag> turns it into profiles/.
set -u
mkdir -p gpurun_out
TAG=${1:-r02z}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; echo "ref rc=$?"
cut -c1-300 gpurun_out/bench_ref_$TAG.json
NCU="ncu --clock-control none"
timeout 900 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -c 1200 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/ncu_list_$TAG.log 2>&1
echo "ncu list rc=$?"
full() {  # name, kernel regex, skip, command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 900 $NCU --set full --import-source on --kernel-name-base demangled -k "regex:$rx" -s $skip -c 1 \
      -f -o gpurun_out/prof_${TAG}_$name "$@" >> gpurun_out/ncu_full_$TAG.log 2>&1
  echo "ncu $name rc=$?"
}
: > gpurun_out/ncu_full_$TAG.log
full block 'tc_up24_block_kernel' 3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra-configs
full up3c2 'tc_conv_kernel<.*6>' 15 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra-configs
full encgemm 'tc_conv_kernel<.*1>' 0 python tools/bench_configs.py --configs 3 --steps 1
full knn 'tc_conv_kernel<.*8>' 0 python tools/bench_configs.py --configs 3 --steps 1
full fft 'stft_fft_kernel' 0 python tools/bench_configs.py --configs 3 --steps 1
python tools/tick_latency.py 1 8 32 128 > gpurun_out/tick_latency_$TAG.log 2>&1; tail -4 gpurun_out/tick_latency_$TAG.log
python tools/profile_tick.py > gpurun_out/profile_tick_$TAG.log 2>&1
python tools/profile_shape.py 64 500 > gpurun_out/profile_c4_$TAG.log 2>&1; head -1 gpurun_out/profile_c4_$TAG.log
timeout 600 $NCU --metrics gpu__time_duration.sum -c 700 --csv --log-file gpurun_out/tick_launches_$TAG.csv python tools/tick_once.py > gpurun_out/ncu_tick_$TAG.log 2>&1
echo "ncu tick rc=$?"
ls -la gpurun_out | tail -12

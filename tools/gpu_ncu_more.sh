#!/bin/bash
TAG=${1:-r02s}
NCU="ncu --clock-control none"
full() {
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 900 $NCU --set full --import-source on --kernel-name-base demangled -k "regex:$rx" -s $skip -c 1 \
      -f -o gpurun_out/prof_${TAG}_$name "$@" >> gpurun_out/ncu_full_$TAG.log 2>&1
  echo "ncu $name rc=$?"; ls gpurun_out/prof_${TAG}_$name.ncu-rep 2>/dev/null
}
full up3c2 'tc_conv_kernel<.*6>' 15 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra-configs
full encgemm 'tc_conv_kernel<.*1>' 0 python tools/bench_configs.py --configs 3 --steps 1
full knn 'tc_conv_kernel<.*8>' 0 python tools/bench_configs.py --configs 3 --steps 1
grep -i "no kernels" gpurun_out/ncu_full_$TAG.log | wc -l

#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r01i}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu_$TAG.log
grep -E "knn_idx|config3" gpurun_out/pytest_gpu_$TAG.log | head
python tools/bench_configs.py --configs 3,5 --steps 3 2>&1 | tail -2
TVC_KNN_EXACT=1 python tools/bench_configs.py --configs 5 --steps 3 2>&1 | tail -1

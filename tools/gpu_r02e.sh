#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc_conv.py -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_r02e.log
bash tools/gpu_ab_opts.sh "pad_max_t=0" "pad_max_t=1024" "pad_max_t=0" "pad_max_t=1024" "pad_max_t=512" "pad_max_t=128" 2>&1 | tee gpurun_out/r02e_pad_ab.log

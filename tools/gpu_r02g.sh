#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused_block.py -x -q 2>&1 | tail -15
FB_CASES="2x1,3x2,64x18,8x500" timeout 600 python tools/fused_block_check.py 2>&1 | tail -6
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_r02g.log
bash tools/gpu_ab_opts.sh "fused_up=0" "" "fused_up=0" "" 2>&1 | tee gpurun_out/r02g_fused_ab.log
python tools/profile_shape.py 64 500 2>&1 | head -12 | tee gpurun_out/r02g_profile_c4.log

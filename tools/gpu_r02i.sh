#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_r02i.log
bash tools/gpu_ab_opts.sh "pdl=0" "" "pdl=0" "" 2>&1 | tee gpurun_out/r02i_ab.log
python tools/trace_run.py 1,14,23,33 gpurun_out/tc_trace_r02i.txt && python tools/trace_view2.py gpurun_out/tc_trace_r02i.txt 24 > gpurun_out/tc_trace_r02i_view.txt

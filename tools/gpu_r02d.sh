#!/bin/bash
# tests + the default bench line (N = 1, all extras) + reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_r02d.log
( time timeout 900 python bench.py > gpurun_out/bench_r02d.json 2> gpurun_out/bench_r02d.err ) 2>&1 | tail -3
tail -c 6000 gpurun_out/bench_r02d.json
tail -5 gpurun_out/bench_r02d.err

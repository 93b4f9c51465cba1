#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/bench_configs.py --configs 3,5 --steps 3 2>&1 | tail -2

#!/bin/bash
# full GPU suite + config-2 bench line (no extras); args: TVC_OPTS strings to A/B ("" = defaults)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
bash tools/gpu_ab_opts.sh "$@"

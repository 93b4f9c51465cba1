#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_ab_opts.sh "" "pad_down_max_t=128" "pad_down_max_t=512" "pad_down_max_t=2047" "" "pad_down_max_t=2047,pad_up_max_t=0" 2>&1 | tee gpurun_out/r02f_pad_ab.log
python - <<'PY'
import json
PY

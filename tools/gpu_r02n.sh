#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encoder_knn.py -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_r02n.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/parity_report.json"))
for k in ("encoder","pipeline_stages","build_index","knn_idx_N50000","config3"):
    if k in d: print(k, d[k])
PY
python tools/profile_tick.py 128 2>&1 | tee gpurun_out/r02n_profile_tick.log
TVC_OPTS="encoder_impl=fp32" python tools/profile_tick.py 128 2>&1 | head -4

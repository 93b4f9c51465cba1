#!/bin/bash
# First GPU call of the next round: validates and times the two experimental paths written blind at the end of round 1.
#   1. fused 24-channel Upsample block (tvc_set_option fused_up=1): bit-identity against the five launches + step time
#   2. tensor copies for edge windows (TVC_TC_EDGE_TMA=1): full GPU parity suite, then a same-box A/B of the bench
# Every mbarrier wait in these kernels traps after ~2 s, so a protocol bug fails the launch instead of hanging the box;
# the timeouts are a second line of defence.
mkdir -p gpurun_out
{
echo "== fused block check"; timeout 300 python tools/fused_block_check.py; echo "rc=$?"
echo "== parity suite with edge tensor copies"; TVC_TC_EDGE_TMA=1 timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
for cfg in "0 0" "1 0" "0 1" "1 1" "0 0"; do
  set -- $cfg
  TVC_TC_EDGE_TMA=$1 TVC_OPTS=fused_up=$2 timeout 300 python bench.py --no-cpu-baseline --steps 40 > gpurun_out/bench_r2ab.json 2>/dev/null
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_r2ab.json")); k = d["roofline"]["per_kernel_ms_per_step"]
    print("edge_tma=$1 fused_up=$2", round(d["ms_per_step"], 4), round(d["value"] / 1e6, 1),
          {n: k.get(n) for n in ("tc_up4_fused", "tc_up4_c2", "tc_up2_c1", "tc_up2_c2", "tc_up1_c2", "tc_down2_c3", "tc_down3_c1")})
except Exception as e:
    print("edge_tma=$1 fused_up=$2 FAILED", e)
PY
done
} 2>&1 | tee gpurun_out/round2_first.log

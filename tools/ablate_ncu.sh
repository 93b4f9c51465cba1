#!/bin/bash
# ncu-timed ablation of the tc_conv roles on one decoder step (kernel durations, not host-polluted event timings)
for d in 0 1 2 4 7; do
  TVC_TC_DBG=$d QB_N=2 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 74 -c 71 --csv --log-file gpurun_out/abl_$d.csv python tools/prof_step.py > /dev/null 2>&1
  echo "== dbg=$d"; python tools/ncu_list.py gpurun_out/abl_$d.csv brief
done

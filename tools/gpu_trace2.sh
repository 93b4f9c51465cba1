#!/bin/bash
mkdir -p gpurun_out
for d in 0 8 16 24 1; do
  TVC_TC_DBG=$d python tools/trace_run.py "43,47" gpurun_out/tc_trace_dbg$d.txt 2>&1 | tail -1
done

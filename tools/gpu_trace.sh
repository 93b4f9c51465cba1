#!/bin/bash
mkdir -p gpurun_out
python tools/trace_run.py "${1:-23,43,44,47}" gpurun_out/tc_trace.txt 2>&1 | tail -3

#!/bin/bash
# compute-sanitizer over every kernel on small shapes (tools/sanitize_small.py): memcheck, racecheck, synccheck
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  TVC_OPTS="graphs=0" timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_small ok|Error|hazard" gpurun_out/sanitizer_$tool.log | sort | uniq -c | head -12
done

#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tests/sanitize_small.py > gpurun_out/sanitize_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_small done|Error|error" gpurun_out/sanitize_$tool.log | head -8
done

#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck synccheck; do
  echo "== $tool"
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitize_$tool.log
  grep -E "ERROR SUMMARY|rc=|sanitize run done|RACECHECK SUMMARY" gpurun_out/sanitize_$tool.log | tail -4
done

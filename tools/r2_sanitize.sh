#!/bin/bash
mkdir -p gpurun_out
( time timeout 200 python tools/sanitize_check.py ) > gpurun_out/r2q_plain.log 2>&1; tail -4 gpurun_out/r2q_plain.log
for tool in memcheck racecheck synccheck; do
  ( time timeout 260 compute-sanitizer --tool $tool --error-exitcode 1 --print-limit 20 python tools/sanitize_check.py ) > gpurun_out/r2q_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_check done|real" gpurun_out/r2q_$tool.log | tail -4
done

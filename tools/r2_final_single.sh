#!/bin/bash
# Round-2 final single-GPU pass: GPU suite, smoke, bench line + launch list, sanitizer walk-through.
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2z_gpu_tests.log 2>&1; tail -4 gpurun_out/r2z_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2z_smoke.log 2>&1; tail -1 gpurun_out/r2z_smoke.log
timeout 900 python bench.py > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err; cut -c1-260 gpurun_out/r2z_bench_n1.json
for tool in memcheck synccheck racecheck; do
  ( time timeout 400 compute-sanitizer --tool $tool --error-exitcode 1 --print-limit 20 python tools/sanitize_check.py ) > gpurun_out/r2z_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_check done|^real" gpurun_out/r2z_$tool.log | tail -3
done

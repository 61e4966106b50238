#!/bin/bash
# Round-end evidence run (1 GPU): tests, bench, ncu launch list of the bench command, ncu full captures.
set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/final_pytest_gpu.log 2>&1; tail -3 gpurun_out/final_pytest_gpu.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/final_clocks.csv & SMI=$!
timeout 600 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; kill $SMI
tail -c 600 gpurun_out/final_bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 400 --csv --log-file gpurun_out/final_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/final_bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:hub_seg_kernel -s 2 -c 1 -o gpurun_out/final_seg_c4 python tools/hv_probe.py c4 0 4 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:hub_cls_kernel -s 1 -c 1 -o gpurun_out/final_cls_dn_c4 python tools/hv_probe.py c4 0 3 dn > /dev/null 2>&1
ls -la gpurun_out/final_*

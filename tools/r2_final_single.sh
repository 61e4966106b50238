#!/bin/bash
# Round-2 single-GPU evidence: new parity tests, per-config record, ncu launch list of the bench command,
# one full capture of the dominant kernel and of the row engine.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_peer_virtual.py -x -q -k "c3_sample or c_dist_calls or heisenberg_xx" 2>&1 | tail -3
python tools/config_bench.py c3 c2 c1 > gpurun_out/r2h_configs.jsonl 2>&1; cut -c1-400 gpurun_out/r2h_configs.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2h_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2h_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hub_seg_kernel -s 3 -c 1 -o gpurun_out/r2h_seg_c4 python tools/prof_eng.py c4 0 full > gpurun_out/r2h_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hub_eng -s 1 -c 1 -o gpurun_out/r2h_eng_dn_c4 python tools/prof_eng.py c4 11 dn > gpurun_out/r2h_ncu2.log 2>&1
ls -la gpurun_out | tail -8

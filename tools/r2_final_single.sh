#!/bin/bash
# Round-2 single-GPU evidence: full GPU suite, contract bench (both arms), engine timings, ncu launch list of
# the bench command, one full capture of the dominant kernel (roofline.traffic) and of the row engine.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2k_tests.log; cat gpurun_out/r2k_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; cut -c1-300 gpurun_out/r2k_bench.json
python tools/eng_bench.py c4 c16 2>&1 | grep -v "^{" > gpurun_out/r2k_eng_bench.txt; cat gpurun_out/r2k_eng_bench.txt
ncu --set full --clock-control none --import-source on -k regex:hub_seg_kernel -s 1 -c 1 -o gpurun_out/r2k_seg_c4 python tools/prof_eng.py c4 0 full > gpurun_out/r2k_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hub_eng -s 1 -c 1 -o gpurun_out/r2k_eng_dn_c4 python tools/prof_eng.py c4 11 dn > gpurun_out/r2k_ncu2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2k_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2k_bench_under_ncu.log 2>&1
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2k_ref.json 2> gpurun_out/r2k_ref.err; cut -c1-1500 gpurun_out/r2k_ref.json

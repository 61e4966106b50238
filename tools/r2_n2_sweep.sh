#!/bin/bash
# 2-GPU sweep of the pull-transpose tile height (C call), 4x4 sector.
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
mkdir -p gpurun_out
for PR in 32 64 128; do
  CMPY_PULL_ROWS=$PR timeout 200 $TR --master-port 2953$((PR % 7)) tools/dist_check.py c4 > gpurun_out/r2i_n2_pull$PR.log 2>&1
  grep -E "L=16.*peer" gpurun_out/r2i_n2_pull$PR.log | sed "s/^/pull_rows=$PR /"
done
timeout 200 $TR --master-port 29541 tools/dist_lanczos.py c4 > gpurun_out/r2i_n2_lanczos.log 2>&1; tail -1 gpurun_out/r2i_n2_lanczos.log | cut -c1-500

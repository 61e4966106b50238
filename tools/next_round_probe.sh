#!/bin/bash
# First GPU call of the next round: validates and times every opt-in path of round 1 in one go.
#   1 GPU :  gpurun --timeout 900 -- 'bash tools/next_round_probe.sh'
#   N GPUs:  gpurun --gpus 2 --timeout 600 -- 'bash tools/next_round_probe.sh multi 2'
# Outputs land in gpurun_out/probe_*.
mkdir -p gpurun_out
if [ "$1" != "multi" ]; then
  # (a) opt-in kernels: partial pull with virtual ranks
  CMPY_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_zz_experimental.py -x -q > gpurun_out/probe_experimental.log 2>&1
  tail -2 gpurun_out/probe_experimental.log
  # (b) the whole parity suite with engine 2 of the class-major kernel as the default engine
  CMPY_CLS_ENGINE=2 timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/probe_suite_engine2.log 2>&1
  tail -2 gpurun_out/probe_suite_engine2.log
  # (c) engine 0 vs engine 2 timings (C4, 16-site chain, 20-site slab)
  timeout 120 python tools/engine_bench.py 20 2>&1 | tail -3
  # (d) C3 (Heisenberg N = 32) with either engine
  for e in 0 2; do CMPY_CLS_ENGINE=$e timeout 200 python tools/config_bench.py c3 2>&1 | tail -1; done
else
  N=${2:-2}
  run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@"; }
  # parity of the sharded H.v (default path), then the bench workload with each overlap scheme
  timeout 300 run tools/dist_check.py 2>&1 | tail -4
  for env in "X=0" "CMPY_PUSH_ORDER=dn_first" "CMPY_PULL_PARTS=2" "CMPY_PULL_PARTS=4" \
             "CMPY_PUSH_ORDER=dn_first CMPY_PULL_PARTS=2" "CMPY_CLS_ENGINE=2"; do
    echo "== $env"
    env $env timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 29512 bench.py --gpus $N --steps 50 --warmup 5 2>&1 | tail -1 | cut -c1-400
    env $env timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 29513 tools/dist_check.py 2>&1 | grep "L=16\|ok" | tail -3
  done
fi

#!/bin/bash
# Round-2 evidence on N GPUs of one box (bounded: every step has its own timeout).
N=${1:-8}; TAG=${2:-r2f}
OUT=gpurun_out/${TAG}_n${N}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > ${OUT}_clocks.csv &
SMI=$!
timeout 240 $TR --master-port 29521 tools/dist_check.py c4 > ${OUT}_dist_check.log 2>&1; echo "dist_check rc=$?"
grep -E "L=1|dist_check ok|phases|rror" ${OUT}_dist_check.log | tail -8
timeout 180 $TR --master-port 29522 tools/dist_lanczos.py c4 > ${OUT}_lanczos_c4.log 2>&1; echo "lanczos c4 rc=$?"; tail -1 ${OUT}_lanczos_c4.log | cut -c1-900
timeout 300 $TR --master-port 29523 bench.py --gpus $N --steps 20 --warmup 5 > ${OUT}_bench.json 2> ${OUT}_bench.err; echo "bench rc=$?"; cut -c1-2500 ${OUT}_bench.json; tail -2 ${OUT}_bench.err | cut -c1-300
if [ "$N" = "8" ]; then
  timeout 400 $TR --master-port 29524 tools/dist_lanczos.py chain20 4.0 20 > ${OUT}_chain20.log 2>&1; echo "chain20 rc=$?"; tail -1 ${OUT}_chain20.log | cut -c1-900
fi
kill $SMI

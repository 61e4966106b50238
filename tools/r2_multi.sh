#!/bin/bash
# Round-2 multi-GPU evidence run: parity of the sharded H.v (C call), sharded Lanczos (C vs Python recurrence),
# overlap options, contract bench.  usage: tools/r2_multi.sh <ngpus> <tag>
N=${1:-2}; TAG=${2:-r2}
OUT=gpurun_out/${TAG}_n${N}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tools/dist_check.py > ${OUT}_dist_check.log 2>&1; echo "dist_check rc=$?"
grep -E "L=16|dist_check ok|phases|Error|error" ${OUT}_dist_check.log | tail -8
CMPY_DIST_PYTHON=1 timeout 600 $TR --master-port 29512 tools/dist_check.py > ${OUT}_dist_check_python.log 2>&1
grep -E "L=16.*peer" ${OUT}_dist_check_python.log | sed 's/^/python-path /'
for opt in "CMPY_PUSH_ORDER=dn_first" "CMPY_PULL_PARTS=2" "CMPY_PULL_PARTS=4"; do
  env CMPY_DIST_PYTHON=1 $opt timeout 600 $TR --master-port 29513 tools/dist_check.py > ${OUT}_opt.log 2>&1
  grep -E "L=16.*peer" ${OUT}_opt.log | sed "s/^/$opt /"
done
timeout 600 $TR --master-port 29514 tools/dist_lanczos.py c4 > ${OUT}_lanczos_c.log 2>&1; tail -1 ${OUT}_lanczos_c.log | cut -c1-700
DIST_LANCZOS_VERBOSE=1 timeout 600 $TR --master-port 29515 tools/dist_lanczos.py c4 > ${OUT}_lanczos_py.log 2>&1; tail -1 ${OUT}_lanczos_py.log | cut -c1-400
timeout 900 $TR --master-port 29516 bench.py --gpus $N --steps 20 --warmup 5 > ${OUT}_bench.json 2> ${OUT}_bench.err; cut -c1-1800 ${OUT}_bench.json; tail -3 ${OUT}_bench.err

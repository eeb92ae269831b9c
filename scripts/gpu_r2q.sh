#!/bin/bash
# Round 2, GPU session Q (8 GPUs, charged 8x: keep it short): sharded parity on 8 ranks (3-qubit swap exchange with the
# block-cyclic partner schedule, paired rounds) and the N = 8 bench line exactly as the driver launches it.
TAG=${1:-r2q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
export QCB_MGC_LOCAL="12:6,17:8,20:8"
echo "== SPMD check N=8"; timeout 400 $TR --nproc-per-node 8 --master-port 29511 tests/multi_gpu_check.py > $OUT/check8.log 2>&1; echo "exit $?"; grep -E "^n=|multi-gpu ok|Error|error" $OUT/check8.log | tail -6
echo "== bench N=8 (as the driver runs it)"
timeout 900 $TR --nproc-per-node 8 --master-port 29515 bench.py --gpus 8 --steps 3 --warmup 2 > $OUT/bench8.log 2>&1; echo "exit $?"; tail -1 $OUT/bench8.log | cut -c1-3000
ls -la $OUT

#!/bin/bash
# Round 2, GPU session O (4 GPUs, charged 4x: keep it short): sharded parity on 4 and 2 ranks with paired rounds and the
# block-cyclic partner schedule of the swap kernel, the single-process handle, the N = 4 bench line as the driver launches it
# (parity block, 35-qubit block, single-process block) and an N = 2 line.
TAG=${1:-r2o}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
scripts/dmma_mix > $OUT/dmma_mix.log 2>&1; tail -5 $OUT/dmma_mix.log | cut -c1-160
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
export QCB_MGC_LOCAL="12:6,17:8,20:8"
echo "== SPMD check N=4"; timeout 600 $TR --nproc-per-node 4 --master-port 29513 tests/multi_gpu_check.py > $OUT/check4.log 2>&1; echo "exit $?"; grep -E "^n=|multi-gpu ok|Error|error" $OUT/check4.log | tail -6
echo "== group check N=4"; QCB_MGC_LOCAL="12:6,18:8,21:8" timeout 600 python tests/group_check.py 4 > $OUT/group4.log 2>&1; echo "exit $?"; tail -4 $OUT/group4.log
echo "== bench N=4 (as the driver runs it)"
timeout 1200 $TR --nproc-per-node 4 --master-port 29515 bench.py --gpus 4 --steps 3 --warmup 2 > $OUT/bench4.log 2>&1; echo "exit $?"; tail -1 $OUT/bench4.log | cut -c1-5000
echo "== SPMD check N=2"; CUDA_VISIBLE_DEVICES=0,1 timeout 600 $TR --nproc-per-node 2 --master-port 29517 tests/multi_gpu_check.py > $OUT/check2.log 2>&1; echo "exit $?"; grep -E "^n=|multi-gpu ok|Error|error" $OUT/check2.log | tail -4
echo "== bench N=2"
CUDA_VISIBLE_DEVICES=0,1 timeout 900 $TR --nproc-per-node 2 --master-port 29519 bench.py --gpus 2 --steps 3 --warmup 2 --no-weak33 --no-single-process > $OUT/bench2.log 2>&1; echo "exit $?"; tail -1 $OUT/bench2.log | cut -c1-3000
ls -la $OUT

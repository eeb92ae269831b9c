#!/bin/bash
# Round 2, GPU session E (8 GPUs, charged 8x: keep it short): sharded parity on 8 and 4 ranks (3- and 2-qubit swap
# exchanges), the single-process handle on 8 GPUs, the full N = 8 bench line (parity block, 33 q / GPU block, single-process
# block) exactly as the driver launches it, and a 4-GPU line.
TAG=${1:-r2e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
nvidia-smi topo -m > $OUT/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
export QCB_MGC_LOCAL="12:6,17:8,20:8"
echo "== SPMD check N=8"; timeout 600 $TR --nproc-per-node 8 --master-port 29511 tests/multi_gpu_check.py > $OUT/check8.log 2>&1; echo "exit $?"; grep -E "^n=|multi-gpu ok|Error|error" $OUT/check8.log | tail -6
echo "== SPMD check N=4"; CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 600 $TR --nproc-per-node 4 --master-port 29513 tests/multi_gpu_check.py > $OUT/check4.log 2>&1; echo "exit $?"; grep -E "^n=|multi-gpu ok|Error|error" $OUT/check4.log | tail -5
echo "== group check N=8"; QCB_MGC_LOCAL="12:6,18:8,21:8" timeout 600 python tests/group_check.py 8 > $OUT/group8.log 2>&1; echo "exit $?"; tail -5 $OUT/group8.log
echo "== bench N=8 (as the driver runs it)"
timeout 1200 $TR --nproc-per-node 8 --master-port 29515 bench.py --gpus 8 --steps 3 --warmup 2 > $OUT/bench8.log 2>&1; echo "exit $?"; tail -1 $OUT/bench8.log | cut -c1-6000
echo "== bench N=4"
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 900 $TR --nproc-per-node 4 --master-port 29517 bench.py --gpus 4 --steps 3 --warmup 2 --no-single-process > $OUT/bench4.log 2>&1; echo "exit $?"; tail -1 $OUT/bench4.log | cut -c1-4000
ls -la $OUT

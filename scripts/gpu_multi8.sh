#!/bin/bash
# 8-GPU session: sharded parity, weak-scaling bench at 30 q/GPU (33 q) and the north-star 36-qubit run (33 q/GPU).
N=${1:-8}; TAG=${2:-multi$N}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > $OUT/check.log 2>&1; echo "check exit $?"; grep "^n=\|multi-gpu ok" $OUT/check.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 1 > $OUT/bench_30q.log 2>&1; echo "bench 30q/GPU exit $?"; tail -1 $OUT/bench_30q.log | cut -c1-2500
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 2 --warmup 1 --qubits 33 > $OUT/bench_33q.log 2>&1; echo "bench 33q/GPU exit $?"; tail -1 $OUT/bench_33q.log | cut -c1-2500

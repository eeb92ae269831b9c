#!/bin/bash
# Multi-GPU session (gpurun --gpus N): sharded parity vs the oracle, then the weak-scaling bench at N ranks.
N=${1:-2}; TAG=${2:-multi$N}; Q=${3:-30}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > $OUT/check.log 2>&1; echo "check exit $?"; tail -6 $OUT/check.log
[ -n "$SKIP_BENCH" ] || timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 1 --qubits $Q > $OUT/bench.log 2>&1; echo "bench exit $?"; tail -2 $OUT/bench.log | cut -c1-3000

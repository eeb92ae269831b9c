#!/bin/bash
# Weak-scaling bench at N ranks (30 qubits per GPU), with the multi-GPU e2e leg.
N=${1:-2}; TAG=${2:-mb$N}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 1 > $OUT/bench.log 2>&1; echo "bench exit $?"; tail -1 $OUT/bench.log | cut -c1-3000

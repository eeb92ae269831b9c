#!/bin/bash
# Round 2, GPU session V (1 GPU): start-up stagger of the two consumer groups on / off; quick parity.
TAG=${1:-r2v}
OUT=gpurun_out/$TAG
mkdir -p $OUT
B="python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e --no-hbm-leg --no-other"
run() { echo "-- $1" | tee -a $OUT/ab.log; shift; env "$@" 2>&1 | tail -1 | python scripts/bench_brief.py | tee -a $OUT/ab.log; }
echo "== parity"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "paired or config3_brickwork or all_gates or variants or generic_ops or grover" 2>&1 | tail -2 | tee $OUT/pytest_quick.log
echo "== A/B"
run "stagger on (default)" X=1 timeout 300 $B
run "stagger off" QCB_STAGGER=0 timeout 300 $B
run "stagger on again" X=1 timeout 300 $B
run "stagger off again" QCB_STAGGER=0 timeout 300 $B
run "stagger on, single rounds r5" QCB_PAIR_ROUNDS=0 timeout 300 $B
run "stagger off, single rounds r5" QCB_STAGGER=0 QCB_PAIR_ROUNDS=0 timeout 300 $B
run "stagger on, stage-rounds 3" X=1 timeout 300 $B --stage-rounds 3
run "stagger off, stage-rounds 3" QCB_STAGGER=0 timeout 300 $B --stage-rounds 3

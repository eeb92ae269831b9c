#!/bin/bash
# Round 2, GPU session U (1 GPU): pass heads fetched inside the previous pass (no dependent table look-ups at the start of a
# pass) against the previous build; quick parity.
TAG=${1:-r2u}
OUT=gpurun_out/$TAG
mkdir -p $OUT
B="python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e --no-hbm-leg --no-other"
run() { echo "-- $1" | tee -a $OUT/ab.log; shift; env "$@" 2>&1 | tail -1 | python scripts/bench_brief.py | tee -a $OUT/ab.log; }
echo "== parity"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "paired or config3_brickwork or all_gates or variants or generic_ops or grover" 2>&1 | tail -2 | tee $OUT/pytest_quick.log
echo "== A/B"
run "pass heads prefetched (new)" X=1 timeout 300 $B
run "previous build" QCB_LIB=qclojure_b200/lib_var/libqcb200_prev.so timeout 300 $B
run "new again" X=1 timeout 300 $B
run "previous again" QCB_LIB=qclojure_b200/lib_var/libqcb200_prev.so timeout 300 $B
run "new, single rounds r5" QCB_PAIR_ROUNDS=0 timeout 300 $B
run "previous, single rounds r5" QCB_LIB=qclojure_b200/lib_var/libqcb200_prev.so QCB_PAIR_ROUNDS=0 timeout 300 $B

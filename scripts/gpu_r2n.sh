#!/bin/bash
# Round 2, GPU session N (1 GPU): operand prefetch inside the pass vs between passes, lockstep vs skewed pipeline of the paired
# rounds, ablations of the tile kernel with results kept alive, quick parity on the variants.
TAG=${1:-r2n}
OUT=gpurun_out/$TAG
mkdir -p $OUT
B="python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e --no-hbm-leg --no-other"
run() { echo "-- $1" | tee -a $OUT/ab.log; shift; env "$@" 2>&1 | tail -1 | python scripts/bench_brief.py | tee -a $OUT/ab.log; }
abl() { echo "-- $1" | tee -a $OUT/ablate.log; shift; env "$@" 2>&1 | grep -E "tile-prof|^\{" | python scripts/ablate_brief.py | tee -a $OUT/ablate.log; }
V=qclojure_b200/lib_var
echo "== parity of the variants"
for lib in qclojure_b200/lib/libqcb200.so $V/libqcb200_lock.so; do
  QCB_LIB=$lib timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "paired or config3_brickwork or all_gates or variants" 2>&1 | tail -2 | tee -a $OUT/pytest_variants.log
done
echo "== A/B"
run "prefetch inside the pass, skewed pipeline (default)" X=1 timeout 300 $B
run "prefetch between passes" QCB_LIB=$V/libqcb200_pfearly.so timeout 300 $B
run "lockstep pairs" QCB_LIB=$V/libqcb200_lock.so timeout 300 $B
run "single rounds r5" QCB_PAIR_ROUNDS=0 timeout 300 $B
run "default again" X=1 timeout 300 $B
run "lockstep pairs again" QCB_LIB=$V/libqcb200_lock.so timeout 300 $B
run "lockstep, eff 150 K4" QCB_LIB=$V/libqcb200_lock.so QCB_PAIR_EFF_PCT=150 QCB_PAIR_SEARCH=4 timeout 300 $B
echo "== ablations (profiling build)"
P=qclojure_b200/lib_prof/libqcb200_ablate.so
BA="python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --no-hbm-leg --no-other"
abl "dbg 0" QCB_LIB=$P QCB_TILE_DBG=0 timeout 300 $BA
abl "dbg 4: no mover HBM traffic" QCB_LIB=$P QCB_TILE_DBG=4 timeout 300 $BA
abl "dbg 2: no fragment LDS/STS" QCB_LIB=$P QCB_TILE_DBG=2 timeout 300 $BA
abl "dbg 6: neither" QCB_LIB=$P QCB_TILE_DBG=6 timeout 300 $BA
abl "dbg 14: neither, no round barriers" QCB_LIB=$P QCB_TILE_DBG=14 timeout 300 $BA
abl "dbg 30: neither, no barriers, A fetched once" QCB_LIB=$P QCB_TILE_DBG=30 timeout 300 $BA
abl "dbg 1: no DMMA" QCB_LIB=$P QCB_TILE_DBG=1 timeout 300 $BA
abl "dbg 6, single rounds r5" QCB_LIB=$P QCB_TILE_DBG=6 QCB_PAIR_ROUNDS=0 timeout 300 $BA
ls -la $OUT

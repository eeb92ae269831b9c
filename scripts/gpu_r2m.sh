#!/bin/bash
# Round 2, GPU session M (1 GPU): paired rounds with the pipelined loop on variant runs and the two-singles comparison rule;
# ablations of the tile kernel with paired rounds (profiling build: no mover traffic / no fragment LDS+STS / no barriers /
# no A prefetch), new GPU tests (paired rounds on the device, swap-kernel unit test, kernel variants).
TAG=${1:-r2m}
OUT=gpurun_out/$TAG
mkdir -p $OUT
B="python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e --no-hbm-leg --no-other"
run() { echo "-- $1" | tee -a $OUT/ab.log; shift; env "$@" 2>&1 | tail -1 | python scripts/bench_brief.py | tee -a $OUT/ab.log; }
abl() { echo "-- $1" | tee -a $OUT/ablate.log; shift; env "$@" 2>&1 | grep -E "tile-prof|^\{" | python scripts/ablate_brief.py | tee -a $OUT/ablate.log; }
echo "== A/B"
run "single rounds r5" QCB_PAIR_ROUNDS=0 timeout 300 $B
run "paired default (budget 7 rounds, pair cost 7/4, eff 170, K 1)" X=1 timeout 300 $B
run "paired r6 cq6" QCB_PAIR_COST_Q=6 timeout 300 $B --stage-rounds 6
run "paired r7 cq7 eff150 K4" QCB_PAIR_EFF_PCT=150 QCB_PAIR_SEARCH=4 timeout 300 $B
run "paired r6 cq7 eff160 K4" QCB_PAIR_EFF_PCT=160 QCB_PAIR_SEARCH=4 timeout 300 $B --stage-rounds 6
run "paired r7 cq6 eff170 K4" QCB_PAIR_COST_Q=6 QCB_PAIR_SEARCH=4 timeout 300 $B
run "paired default, consumers 2x8" QCB_CONSUMERS=2x8 timeout 300 $B
echo "== ablations (profiling build, paired default)"
P=qclojure_b200/lib_prof/libqcb200_ablate.so
BA="python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --no-hbm-leg --no-other"
abl "dbg 0" QCB_LIB=$P QCB_TILE_DBG=0 timeout 300 $BA
abl "dbg 4: no mover HBM traffic" QCB_LIB=$P QCB_TILE_DBG=4 timeout 300 $BA
abl "dbg 2: no fragment LDS/STS" QCB_LIB=$P QCB_TILE_DBG=2 timeout 300 $BA
abl "dbg 6: neither" QCB_LIB=$P QCB_TILE_DBG=6 timeout 300 $BA
abl "dbg 14: neither, no round barriers" QCB_LIB=$P QCB_TILE_DBG=14 timeout 300 $BA
abl "dbg 30: neither, no barriers, A fetched once" QCB_LIB=$P QCB_TILE_DBG=30 timeout 300 $BA
abl "dbg 1: no DMMA" QCB_LIB=$P QCB_TILE_DBG=1 timeout 300 $BA
abl "dbg 5: no DMMA, no mover traffic" QCB_LIB=$P QCB_TILE_DBG=5 timeout 300 $BA
abl "dbg 0, single rounds r5" QCB_LIB=$P QCB_TILE_DBG=0 QCB_PAIR_ROUNDS=0 timeout 300 $BA
abl "dbg 6, single rounds r5" QCB_LIB=$P QCB_TILE_DBG=6 QCB_PAIR_ROUNDS=0 timeout 300 $BA
echo "== new GPU tests"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "paired or swap_kernel or variants" > $OUT/pytest_gpu_new.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu_new.log
tail -4 $OUT/pytest_gpu_new.log
ls -la $OUT

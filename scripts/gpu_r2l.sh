#!/bin/bash
# Round 2, GPU session L (1 GPU): paired rounds (round kind 3) - A/B against single rounds, budget / yield knobs, the GPU
# test-suite incl. the 30-qubit comparison with the C oracle, one full ncu capture of the tile kernel.
TAG=${1:-r2l}
OUT=gpurun_out/$TAG
mkdir -p $OUT
B="python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e --no-hbm-leg --no-other"
run() { echo "-- $1" | tee -a $OUT/ab.log; shift; env "$@" 2>&1 | tail -1 | python scripts/bench_brief.py | tee -a $OUT/ab.log; }
echo "== A/B"
run "single rounds r5 (round-2 default so far)" QCB_PAIR_ROUNDS=0 timeout 300 $B
run "paired rounds, default budget (20 quarter rounds, pair = 6, yield 60)" X=1 timeout 300 $B
run "paired, stage-rounds 6" X=1 timeout 300 $B --stage-rounds 6
run "paired, stage-rounds 7" X=1 timeout 300 $B --stage-rounds 7
run "paired, stage-rounds 8" X=1 timeout 300 $B --stage-rounds 8
run "paired, yield 40" QCB_PAIR_YIELD_PCT=40 timeout 300 $B
run "paired, yield 80" QCB_PAIR_YIELD_PCT=80 timeout 300 $B
run "paired, yield 100" QCB_PAIR_YIELD_PCT=100 timeout 300 $B
run "paired, pair cost 5" QCB_PAIR_COST_Q=5 timeout 300 $B
run "paired, pair cost 7, stage-rounds 7" QCB_PAIR_COST_Q=7 timeout 300 $B --stage-rounds 7
run "paired, stage-rounds 3" X=1 timeout 300 $B --stage-rounds 3
run "paired, stage-rounds 4" X=1 timeout 300 $B --stage-rounds 4
run "single rounds r3" QCB_PAIR_ROUNDS=0 timeout 300 $B --stage-rounds 3
echo "== pytest gpu, everything incl. the 30 q oracle comparison"
timeout 2400 python -m pytest tests -m gpu -q -x --timeout 1800 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
echo "== ncu full, 28 qubits"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_stage -s 20 -c 2 -o $OUT/prof_tile \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-hbm-leg --no-other --qubits 28 > $OUT/ncu_full.log 2>&1; echo "ncu full exit $?"; tail -2 $OUT/ncu_full.log | cut -c1-300
ls -la $OUT

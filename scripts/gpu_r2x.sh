#!/bin/bash
# Round 2, GPU session X (1 GPU): far phases (QFT ladders) - parity, headline unchanged?, QFT timings with / without.
TAG=${1:-r2x}
OUT=gpurun_out/$TAG
mkdir -p $OUT
B="python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e --no-hbm-leg --no-other"
run() { echo "-- $1" | tee -a $OUT/ab.log; shift; env "$@" 2>&1 | tail -1 | python scripts/bench_brief.py | tee -a $OUT/ab.log; }
echo "== parity"
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -k "not spot_amplitudes_vs_c_oracle" > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
echo "== headline"
run "default" X=1 timeout 300 $B
run "default again" X=1 timeout 300 $B
echo "== QFT"
timeout 600 python scripts/qft_probe.py 2>&1 | tee $OUT/qft_probe.log

#!/bin/bash
# Round 2, GPU session C (1 GPU): direct store A/B, budgets, full GPU test-suite, ncu of the new hot kernel, full bench line.
TAG=${1:-r2c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
B="python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --no-hbm-leg --no-other"
run() { echo "-- $1" | tee -a $OUT/ab.log; shift; env "$@" 2>&1 | tail -1 | python scripts/bench_brief.py | tee -a $OUT/ab.log; }
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -k "not spot_amplitudes_vs_c_oracle or 28" > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
echo "== A/B"
run "direct store r5 (default)" X=1 timeout 300 $B
run "no direct store r5" QCB_DIRECT_STORE=0 timeout 300 $B
run "direct store, rotating loop r5" QCB_LIB=qclojure_b200/lib_var/libqcb200_rotate.so timeout 300 $B
for r in 2 3 4 6 8; do run "direct stage-rounds $r" X=1 timeout 300 $B --stage-rounds $r; done
run "no direct stage-rounds 3" QCB_DIRECT_STORE=0 timeout 300 $B --stage-rounds 3
run "direct, 4 buffers? (QCB_TILE_BUFFERS=3 is the max that fits)" QCB_TILE_BUFFERS=3 timeout 300 $B --stage-rounds 3
echo "== full bench line"
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.log 2>&1; echo "bench exit $?"; tail -1 $OUT/bench.log | cut -c1-3000
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.log 2>&1; tail -1 $OUT/bench_reference.log | cut -c1-1200
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-hbm-leg --no-other > $OUT/ncu_launches.log 2>&1; echo "ncu launches exit $?"
echo "== ncu full (30 qubits, one 5-round launch)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_stage -s 20 -c 1 -o $OUT/prof_tile_30q \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-hbm-leg --no-other > $OUT/ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la $OUT

#!/bin/bash
# Round 2, GPU session A (1 GPU): parity of the three-product tensor-core rounds, A/B of kernel variants, budgets, ncu.
TAG=${1:-r2a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest gpu (without the 30 q oracle comparison)"
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 -k "not spot_amplitudes_vs_c_oracle or 28" > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
B="python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --no-hbm-leg"
run() { echo "-- $1" | tee -a $OUT/ab.log; shift; env "$@" 2>&1 | tail -1 | python scripts/bench_brief.py | tee -a $OUT/ab.log; }
echo "== A/B"
run "legacy m16n8k16 form r5" QCB_MMA_FORM=1 timeout 300 $B
run "k3 rolled r5 (default)" X=1 timeout 300 $B
run "k3 unrolled r5" QCB_LIB=qclojure_b200/lib_var/libqcb200_unroll.so timeout 300 $B
run "k3 realloc r5" QCB_LIB=qclojure_b200/lib_var/libqcb200_realloc.so timeout 300 $B
run "k3 consumers 1x8 r5" QCB_CONSUMERS=1x8 timeout 300 $B
for r in 2 3 4 6 8; do
  run "k3 rolled stage-rounds $r" X=1 timeout 300 $B --stage-rounds $r
done
run "legacy stage-rounds 3" QCB_MMA_FORM=1 timeout 300 $B --stage-rounds 3
run "k3 tile-bits 11 stage-rounds 3" X=1 timeout 300 $B --tile-bits 11 --stage-rounds 3
run "k3 tile-bits 11 stage-rounds 5" X=1 timeout 300 $B --tile-bits 11
run "k3 unrolled stage-rounds 3" QCB_LIB=qclojure_b200/lib_var/libqcb200_unroll.so timeout 300 $B --stage-rounds 3
echo "== full bench line"
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.log 2>&1; echo "bench exit $?"; tail -1 $OUT/bench.log | cut -c1-2500
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-hbm-leg > $OUT/ncu_launches.log 2>&1; echo "ncu launches exit $?"
echo "== ncu full (30 qubits, one 5-round launch)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_stage -s 20 -c 1 -o $OUT/prof_tile_30q \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-hbm-leg > $OUT/ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la $OUT

#!/bin/bash
# Round 2, GPU session G (1 GPU): loop-unroll / ptxas variants, reductions after the quad-parity change, racecheck on a
# single-tile run, the 30-qubit comparison with the C oracle, full bench line.
TAG=${1:-r2g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
B="python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --no-hbm-leg --no-other"
run() { echo "-- $1" | tee -a $OUT/ab.log; shift; env "$@" 2>&1 | tail -1 | python scripts/bench_brief.py | tee -a $OUT/ab.log; }
echo "== swap kernel test"; timeout 300 tests/cuda/swap_kernel_test 2>&1 | tee $OUT/swap_kernel_test.log
echo "== A/B"
run "default r5" X=1 timeout 300 $B
run "pp loop unroll 2 r5" QCB_LIB=qclojure_b200/lib_var/libqcb200_unroll2.so timeout 300 $B
run "ptxas -O2 r5" QCB_LIB=qclojure_b200/lib_var/libqcb200_o2.so timeout 300 $B
run "default r3" X=1 timeout 300 $B --stage-rounds 3
run "pp loop unroll 2 r3" QCB_LIB=qclojure_b200/lib_var/libqcb200_unroll2.so timeout 300 $B --stage-rounds 3
run "default r4" X=1 timeout 300 $B --stage-rounds 4
echo "== racecheck, single tile"
timeout 900 compute-sanitizer --tool racecheck --print-limit 40 python tests/sanitize_check.py --single-tile > $OUT/racecheck_single_tile.log 2>&1; echo "exit $?"
grep -E "^ok |sanitize_check ok|SUMMARY" $OUT/racecheck_single_tile.log; grep -E "Error: Race|Warning: Race|and (Read|Write) access" $OUT/racecheck_single_tile.log | sed 's/+0x[0-9a-f]*//g; s/\[[0-9]* hazards\]//' | sort | uniq -c | sort -rn | head
echo "== reductions"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'k_reduce|k_expect|k_chunk|k_scale|k_prob|k_marg|k_grover' -c 20 --csv --log-file $OUT/reductions.csv \
    python scripts/reduction_probe.py > $OUT/reduction_probe.log 2>&1; echo "ncu reductions exit $?"
echo "== pytest gpu, everything incl. the 30 q oracle comparison"
timeout 2400 python -m pytest tests -m gpu -q -x --timeout 1800 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
echo "== full bench line"
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.log 2>&1; echo "bench exit $?"; tail -1 $OUT/bench.log | cut -c1-1500
ls -la $OUT

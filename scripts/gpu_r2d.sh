#!/bin/bash
# Round 2, GPU session D (1 GPU): consumer layout 2x8 probe, new reduction / measurement code, sanitizer runs.
TAG=${1:-r2d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
B="python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --no-hbm-leg --no-other"
run() { echo "-- $1" | tee -a $OUT/ab.log; shift; env "$@" 2>&1 | tail -1 | python scripts/bench_brief.py | tee -a $OUT/ab.log; }
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -k "not spot_amplitudes_vs_c_oracle" > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
echo "== A/B"
run "2x4 r5 (default)" X=1 timeout 300 $B
run "2x8 r5" QCB_CONSUMERS=2x8 timeout 300 $B
run "2x4 r3" X=1 timeout 300 $B --stage-rounds 3
run "2x8 r3" QCB_CONSUMERS=2x8 timeout 300 $B --stage-rounds 3
run "2x8 r4" QCB_CONSUMERS=2x8 timeout 300 $B --stage-rounds 4
run "2x8 r2" QCB_CONSUMERS=2x8 timeout 300 $B --stage-rounds 2
run "2x8 r8" QCB_CONSUMERS=2x8 timeout 300 $B --stage-rounds 8
echo "== other configs (reductions)"
timeout 600 python scripts/bench_configs.py > $OUT/other_configs.json 2>&1; tail -14 $OUT/other_configs.json
echo "== ncu of the streaming reductions (28 q: norm, probabilities, expectation)"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'k_reduce|k_expect|k_chunk|k_scale|k_prob|k_marg|k_grover' -c 40 --csv --log-file $OUT/reductions.csv \
    python scripts/reduction_probe.py > $OUT/reduction_probe.log 2>&1; echo "ncu reductions exit $?"; tail -3 $OUT/reduction_probe.log
echo "== compute-sanitizer"
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python tests/sanitize_check.py > $OUT/sanitizer_$tool.log 2>&1; echo "$tool exit $?"; tail -4 $OUT/sanitizer_$tool.log
done
ls -la $OUT

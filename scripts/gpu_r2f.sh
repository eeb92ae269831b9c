#!/bin/bash
# Round 2, GPU session F (1 GPU): swap-kernel unit test (k = 1..3 on one device), sanitizer runs, racecheck mbarrier probe,
# reduction bandwidths after the POPC / marginal changes.
TAG=${1:-r2f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== swap kernel test"; timeout 300 tests/cuda/swap_kernel_test 2>&1 | tee $OUT/swap_kernel_test.log
echo "== pytest gpu (measure / marginal / expectation subset)"
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "measure or marginal or expectation or partial or result_extraction or hardware or noisy or qaoa or vqe" > $OUT/pytest_subset.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_subset.log
echo "== reductions"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'k_reduce|k_expect|k_chunk|k_scale|k_prob|k_marg|k_grover' -c 40 --csv --log-file $OUT/reductions.csv \
    python scripts/reduction_probe.py > $OUT/reduction_probe.log 2>&1; echo "ncu reductions exit $?"
echo "== racecheck mbarrier probe"
timeout 300 compute-sanitizer --tool racecheck tests/cuda/mbar_racecheck_probe > $OUT/racecheck_mbar_probe.log 2>&1; tail -12 $OUT/racecheck_mbar_probe.log | cut -c1-200
echo "== compute-sanitizer on every kernel variant"
for tool in memcheck synccheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 40 python tests/sanitize_check.py > $OUT/sanitizer_$tool.log 2>&1; echo "$tool exit $?"; grep -E "^ok |sanitize_check ok|SUMMARY" $OUT/sanitizer_$tool.log | tail -14
done
ls -la $OUT

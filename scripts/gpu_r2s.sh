#!/bin/bash
# Round 2, GPU session S (1 GPU): the state the round ends on after the reduction kernels changed - full GPU test-suite, smoke,
# the full bench line (ours + reference arm), ncu launch list.
TAG=${1:-r2s}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest gpu, everything incl. the 30 q oracle comparison"
timeout 2400 python -m pytest tests -m gpu -q -x --timeout 1800 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/smoke.log; tail -2 $OUT/smoke.log
echo "== full bench line"
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.log 2>&1; echo "bench exit $?"; tail -1 $OUT/bench.log | cut -c1-1500
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.log 2>&1; tail -1 $OUT/bench_reference.log | cut -c1-400
echo "== sanitizer: memcheck on the small-size sweep of every kernel variant"
timeout 900 compute-sanitizer --tool memcheck python tests/sanitize_check.py > $OUT/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?"; grep -E "ERROR SUMMARY|sanitize_check ok" $OUT/sanitizer_memcheck.log | tail -3
ls -la $OUT

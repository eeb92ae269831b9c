#!/bin/bash
# Round 2, last GPU session (1 GPU, ~10 GPU-minutes left): the tree with the 16-per-thread chunk-sum scan and the plan-replay
# lane-role hints - the full bench line first (reductions block: "1024 shots" now without the 1024-step scan), then the whole
# GPU test-suite (incl. the new 25-qubit sampling case that crosses scan tiles).
TAG=${1:-r2zz}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== full bench line"
timeout 240 python bench.py --steps 5 --warmup 3 > $OUT/bench.log 2>&1; echo "bench exit $?"; tail -1 $OUT/bench.log | cut -c1-400
tail -1 $OUT/bench.log | python scripts/bench_brief.py   # (the session ran this with the file as an ARGUMENT: bench_brief reads stdin, so it sat there until the box limit and the test-suite never started)
echo "== pytest gpu (everything)"
timeout 420 python -m pytest tests -m gpu -q -x --timeout 400 --durations=12 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -25 $OUT/pytest_gpu.log

#!/bin/bash
# Round 2, GPU session Z (1 GPU): the tree the round ends on - full GPU test-suite incl. the 30-qubit comparison with the C oracle,
# smoke, the full bench line (ours + reference arm), ncu launch list and one full capture of the tile kernel.
TAG=${1:-r2z}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest gpu (everything)"
timeout 2400 python -m pytest tests -m gpu -q -x --timeout 1800 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/smoke.log; tail -2 $OUT/smoke.log
echo "== full bench line"
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.log 2>&1; echo "bench exit $?"; tail -1 $OUT/bench.log | cut -c1-1200
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.log 2>&1; tail -1 $OUT/bench_reference.log | cut -c1-300
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-hbm-leg --no-other > $OUT/ncu_launches.log 2>&1; echo "ncu launches exit $?"
echo "== ncu full (30 qubits, one launch of the tile kernel)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_stage -s 16 -c 1 -o $OUT/prof_tile_30q \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-hbm-leg --no-other > $OUT/ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la $OUT

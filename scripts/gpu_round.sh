#!/bin/bash
# One GPU session: parity tests, smoke, bench (ours + reference arm), ncu launch list + one full capture of the top kernel.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_round.sh [tag]
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > $OUT/gpu.txt 2>&1
free -g > $OUT/host.txt; nproc >> $OUT/host.txt
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/smoke.log; tail -3 $OUT/smoke.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.log 2>&1; echo "bench exit $?" | tee -a $OUT/bench.log; tail -3 $OUT/bench.log | cut -c1-3000
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.log 2>&1; tail -2 $OUT/bench_reference.log | cut -c1-1500
echo "== bench unfused"; timeout 600 python bench.py --steps 1 --warmup 1 --fusion 0 --no-cpu --no-e2e > $OUT/bench_unfused.log 2>&1; tail -2 $OUT/bench_unfused.log | cut -c1-1500
echo "== bench budget sweep"
for cfg in 200,2 800,3 800,5 800,8 800,12; do
  IFS=, read c r <<< "$cfg"
  echo "-- stage-cost $c stage-rounds $r" >> $OUT/budgets.log
  timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu --no-e2e --stage-cost $c --stage-rounds $r 2>&1 | tail -1 >> $OUT/budgets.log
done
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > $OUT/ncu_launches.log 2>&1; echo "ncu launches exit $?"
echo "== ncu full (30 qubits, one launch: dram traffic per launch)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_stage -s 12 -c 1 -o $OUT/prof_tile_30q \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > $OUT/ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la $OUT

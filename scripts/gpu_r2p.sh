#!/bin/bash
# Round 2, GPU session P (1 GPU): the state the round ends on - full GPU test-suite (incl. the 30-qubit comparison with the C
# oracle), smoke, the full bench line (ours + reference arm), plan portfolio on / off, ncu launch list and one full capture
# of the tile kernel at 30 qubits.
TAG=${1:-r2p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu.txt 2>&1
B="python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e --no-hbm-leg --no-other"
run() { echo "-- $1" | tee -a $OUT/ab.log; shift; env "$@" 2>&1 | tail -1 | python scripts/bench_brief.py | tee -a $OUT/ab.log; }
echo "== A/B"
run "default (paired rounds, plan portfolio)" X=1 timeout 300 $B
run "no portfolio (budget 7 rounds, pair cost 7/4, eff 170, K 1)" QCB_PLAN_PORTFOLIO=0 timeout 300 $B
run "single rounds r5 (the kernel of the round's first sessions)" QCB_PAIR_ROUNDS=0 timeout 300 $B
run "default, stage-rounds 3" X=1 timeout 300 $B --stage-rounds 3
run "default, stage-rounds 2" X=1 timeout 300 $B --stage-rounds 2
run "default again" X=1 timeout 300 $B
echo "== pytest gpu, everything incl. the 30 q oracle comparison"
timeout 2400 python -m pytest tests -m gpu -q -x --timeout 1800 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/smoke.log; tail -2 $OUT/smoke.log
echo "== full bench line"
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.log 2>&1; echo "bench exit $?"; tail -1 $OUT/bench.log | cut -c1-2500
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.log 2>&1; tail -1 $OUT/bench_reference.log | cut -c1-800
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-hbm-leg --no-other > $OUT/ncu_launches.log 2>&1; echo "ncu launches exit $?"
echo "== ncu full (30 qubits, two launches of the tile kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_stage -s 16 -c 2 -o $OUT/prof_tile_30q \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-hbm-leg --no-other > $OUT/ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la $OUT

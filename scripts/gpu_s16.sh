#!/bin/bash
TAG=${1:-s16}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 120 scripts/tma_rate > $OUT/tma_rate.log 2>&1; tail -3 $OUT/tma_rate.log
summ() { tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({k:d[k] for k in ('value','ms_per_step','sweeps_per_step','rounds_per_step','gates_per_sweep')}), json.dumps({k:d['roofline'][k] for k in ('achieved','frac','avg_launch_ms')}), d['roofline']['fp64_tensor']['frac'], d['clocks'])"; }
for lt in 1x8,12 3x4,11 2x4,11; do
IFS=, read lay tb <<< "$lt"
for cfg in 200,4 800,12; do
  IFS=, read c r <<< "$cfg"
  echo "== consumers $lay tile-bits $tb stage-cost $c stage-rounds $r" | tee -a $OUT/sweep.log
  QCB_CONSUMERS=$lay timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu --no-e2e --tile-bits $tb --stage-cost $c --stage-rounds $r 2>&1 | summ | tee -a $OUT/sweep.log
done; done
for q in 31 32 33; do
  echo "== single GPU $q qubits" | tee -a $OUT/big.log
  timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --qubits $q 2>&1 | tail -1 | tee -a $OUT/big.log | cut -c1-400
done

#!/bin/bash
TAG=${1:-s13}
OUT=gpurun_out/$TAG
mkdir -p $OUT
summ() { grep '^{' | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({k:d[k] for k in ('value','ms_per_step','sweeps_per_step','rounds_per_step')}), json.dumps({k:d['roofline'][k] for k in ('frac','avg_launch_ms')}))"; }
for cfg in 200,4 200,2 800,12 1600,24; do
  IFS=, read c r <<< "$cfg"
  echo "== normal build stage-cost $c stage-rounds $r" | tee -a $OUT/prof.log
  timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu --no-e2e --stage-cost $c --stage-rounds $r 2>/dev/null | summ | tee -a $OUT/prof.log
  echo "== PROFILE build stage-cost $c stage-rounds $r" | tee -a $OUT/prof.log
  QCB_LIB=$PWD/qclojure_b200/lib_prof/libqcb200.so timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --stage-cost $c --stage-rounds $r > $OUT/last.log 2>&1; summ < $OUT/last.log | tee -a $OUT/prof.log; grep tile-prof $OUT/last.log | tee -a $OUT/prof.log
done

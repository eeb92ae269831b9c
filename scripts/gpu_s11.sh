#!/bin/bash
# Ablations of the tile kernel (profiling build): which resource sets the sweep time?
TAG=${1:-s11}
OUT=gpurun_out/$TAG
mkdir -p $OUT
summ() { grep '^{' | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({k:d[k] for k in ('ms_per_step','sweeps_per_step','rounds_per_step')}), json.dumps({k:d['roofline'][k] for k in ('frac','avg_launch_ms')}))"; }
for cfg in 200,4 200,2; do
for dbg in 0 1 2 3 4 5 6 7; do
  IFS=, read c r <<< "$cfg"
  echo "== DBG $dbg (1=no DMMA 2=no LDS/STS 4=no HBM) stage-rounds $r" | tee -a $OUT/ablate.log
  QCB_TILE_DBG=$dbg QCB_LIB=$PWD/qclojure_b200/lib_prof/libqcb200.so timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --stage-cost $c --stage-rounds $r > $OUT/last.log 2>&1; summ < $OUT/last.log | tee -a $OUT/ablate.log; [ $dbg = 0 ] && grep tile-prof $OUT/last.log | tee -a $OUT/ablate.log
done; done

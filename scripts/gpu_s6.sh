#!/bin/bash
TAG=${1:-s6}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 120 scripts/dmma_bench2 | tee $OUT/dmma2.log
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/smoke.log
summ() { tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({k:d[k] for k in ('value','ms_per_step','sweeps_per_step','rounds_per_step','gates_per_sweep')}), json.dumps({k:d['roofline'][k] for k in ('achieved','frac','avg_launch_ms')}), d['clocks'])"; }
for lay in ${LAYOUTS:-2x8 2x4}; do
for tb in ${TILEBITS:-11}; do
for cfg in ${SWEEP:-200,3 200,4 400,6}; do
  IFS=, read c r <<< "$cfg"
  echo "== consumers $lay tile-bits $tb stage-cost $c stage-rounds $r" | tee -a $OUT/sweep.log
  QCB_CONSUMERS=$lay timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu --no-e2e --tile-bits $tb --stage-cost $c --stage-rounds $r 2>&1 | summ | tee -a $OUT/sweep.log
done; done; done

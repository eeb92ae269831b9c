#!/bin/bash
TAG=${1:-s17}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
summ() { tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({k:d[k] for k in ('value','ms_per_step','sweeps_per_step','rounds_per_step')}), json.dumps({k:d['roofline'][k] for k in ('frac','avg_launch_ms')}), round(d['roofline']['fp64_tensor']['frac'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
for pz in 0 100 300 1000; do
for cfg in 200,2 800,12; do
  IFS=, read c r <<< "$cfg"
  echo "== pause $pz stage-cost $c stage-rounds $r" | tee -a $OUT/sweep.log
  QCB_MOVER_PAUSE_NS=$pz timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu --no-e2e --stage-cost $c --stage-rounds $r 2>&1 | summ | tee -a $OUT/sweep.log
done; done

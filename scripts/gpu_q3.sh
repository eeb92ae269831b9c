#!/bin/bash
OUT=gpurun_out/${1:-q3}; mkdir -p $OUT
summ() { tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({k:d[k] for k in ('value','ms_per_step','sweeps_per_step','rounds_per_step')}), json.dumps({k:d['roofline'][k] for k in ('frac','avg_launch_ms')}), round(d['roofline']['fp64_tensor']['frac'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
for cfg in 5,50 8,50 12,50 8,0 4,50 3,50; do
  IFS=, read r y <<< "$cfg"
  echo "== max_rounds $r yield $y" | tee -a $OUT/sweep.log
  QCB_ROUND_YIELD_PCT=$y timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu --no-e2e --stage-cost 800 --stage-rounds $r 2>&1 | summ | tee -a $OUT/sweep.log
done

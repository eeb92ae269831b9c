#!/bin/bash
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_gpu.log
summ() { tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({k:d[k] for k in ('value','ms_per_step','sweeps_per_step','rounds_per_step')}), json.dumps({k:d['roofline'][k] for k in ('frac','avg_launch_ms')}), round(d['roofline']['fp64_tensor']['frac'],3), d.get('e2e',{}).get('value'), d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
echo "== default" | tee -a $OUT/sweep.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | summ | tee -a $OUT/sweep.log
for cfg in ${SWEEP:-200,2 200,4}; do
  IFS=, read c r <<< "$cfg"
  echo "== stage-cost $c stage-rounds $r" | tee -a $OUT/sweep.log
  timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu --no-e2e --stage-cost $c --stage-rounds $r 2>&1 | summ | tee -a $OUT/sweep.log
done

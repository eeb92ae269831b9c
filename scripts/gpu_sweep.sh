#!/bin/bash
# Parameter sweep of the fused executor on the 30-qubit benchmark circuit (scheduler budgets, tile geometry).
TAG=${1:-sweep}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_gpu.log
QCB_DENSE_MMA=2 timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > $OUT/pytest_gpu_interp.log 2>&1; echo "pytest (interpreter only) exit $?"; tail -3 $OUT/pytest_gpu_interp.log
[ -x scripts/dmma_bench ] && timeout 120 scripts/dmma_bench | tee $OUT/dmma.log
SWEEP=${SWEEP:-0,0,0,0 200,2,0,0 200,3,0,0 200,4,0,0 200,5,0,0 400,6,0,0 200,3,11,0 200,4,12,6}
for cfg in $SWEEP; do
  IFS=, read c r t l <<< "$cfg"
  echo "== stage-cost $c stage-rounds $r tile-bits $t low-bits $l"
  timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu --no-e2e --stage-cost $c --stage-rounds $r --tile-bits $t --low-bits $l 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({k:d[k] for k in ('value','ms_per_step','sweeps_per_step','rounds_per_step','gates_per_sweep')}), json.dumps({k:d['roofline'][k] for k in ('achieved','frac','avg_launch_ms')}))" | tee -a $OUT/sweep.log
done

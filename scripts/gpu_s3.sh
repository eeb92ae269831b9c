#!/bin/bash
# Pipelined tile kernel: smoke, parity, bench for 16 / 8 consumer warps, budget sweep, one ncu capture.
TAG=${1:-s3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/smoke.log
grep -q "smoke ok" $OUT/smoke.log || { echo "smoke failed, stopping"; exit 1; }
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_gpu.log
summ() { tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({k:d[k] for k in ('value','ms_per_step','sweeps_per_step','rounds_per_step','gates_per_sweep')}), json.dumps({k:d['roofline'][k] for k in ('achieved','frac','avg_launch_ms')}), d['clocks'])"; }
for ncw in 16 8; do
for cfg in ${SWEEP:-0,0 200,2 200,3 200,4 400,6}; do
  IFS=, read c r <<< "$cfg"
  echo "== consumer-warps $ncw stage-cost $c stage-rounds $r" | tee -a $OUT/sweep.log
  QCB_CONSUMER_WARPS=$ncw timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu --no-e2e --stage-cost $c --stage-rounds $r 2>&1 | summ | tee -a $OUT/sweep.log
done; done
echo "== unfused" | tee -a $OUT/sweep.log
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --fusion 0 2>&1 | summ | tee -a $OUT/sweep.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_stage -s 10 -c 2 -o $OUT/prof_tile \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --qubits 28 ${NCU_ARGS:-} > $OUT/ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la $OUT

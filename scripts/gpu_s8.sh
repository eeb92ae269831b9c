#!/bin/bash
TAG=${1:-s8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/smoke.log
grep -q "smoke ok" $OUT/smoke.log || { echo "smoke failed, stopping"; tail -20 $OUT/smoke.log; exit 1; }
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/pytest_gpu.log
summ() { tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({k:d[k] for k in ('value','ms_per_step','sweeps_per_step','rounds_per_step','gates_per_sweep')}), json.dumps({k:d['roofline'][k] for k in ('achieved','frac','avg_launch_ms')}), d['clocks'])"; }
for lay in ${LAYOUTS:-2x4 2x8}; do
for cfg in ${SWEEP:-200,2 200,3 200,4 400,6}; do
  IFS=, read c r <<< "$cfg"
  echo "== consumers $lay stage-cost $c stage-rounds $r" | tee -a $OUT/sweep.log
  QCB_CONSUMERS=$lay timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu --no-e2e --stage-cost $c --stage-rounds $r 2>&1 | summ | tee -a $OUT/sweep.log
done; done
echo "== unfused" | tee -a $OUT/sweep.log
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --fusion 0 2>&1 | summ | tee -a $OUT/sweep.log
for cfg in 200,4 200,2; do
  IFS=, read c r <<< "$cfg"
  echo "== PROFILE stage-cost $c stage-rounds $r" | tee -a $OUT/prof.log
  QCB_LIB=$PWD/qclojure_b200/lib_prof/libqcb200.so timeout 300 python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --stage-cost $c --stage-rounds $r 2>&1 | grep "tile-prof" | tee -a $OUT/prof.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_stage -s 10 -c 2 -o $OUT/prof_tile \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --qubits 28 ${NCU_ARGS:-} > $OUT/ncu_full.log 2>&1; echo "ncu full exit $?"

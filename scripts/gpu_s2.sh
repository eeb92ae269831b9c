#!/bin/bash
# Session 2 measurement: microbench, parity, bench in both executor modes, budget sweep, ncu captures.
TAG=${1:-s2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/host.txt; free -g >> $OUT/host.txt
[ -x scripts/dmma_bench ] && timeout 120 scripts/dmma_bench | tee $OUT/dmma.log
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_gpu.log
summ() { tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({k:d[k] for k in ('value','ms_per_step','sweeps_per_step','rounds_per_step','gates_per_sweep')}), json.dumps({k:d['roofline'][k] for k in ('achieved','frac','avg_launch_ms')}), d['clocks'])"; }
for mode in 1 2; do
for cfg in 0,0 200,2 200,3 200,4 400,6; do
  IFS=, read c r <<< "$cfg"
  echo "== dense-mma $mode stage-cost $c stage-rounds $r" | tee -a $OUT/sweep.log
  timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu --no-e2e --dense-mma $mode --stage-cost $c --stage-rounds $r 2>&1 | summ | tee -a $OUT/sweep.log
done; done
for mode in 1 2; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_stage -s 10 -c 2 -o $OUT/prof_tile_m$mode \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --qubits 28 --dense-mma $mode > $OUT/ncu_full_m$mode.log 2>&1; echo "ncu full exit $?"
done
ls -la $OUT

#!/bin/bash
# Round 2, GPU session W (1 GPU): the final tree - quick parity, bench line, ncu durations / DRAM bytes of the streaming reductions.
TAG=${1:-r2w}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== reductions under ncu (28 qubits)"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'k_reduce|k_expect|k_chunk|k_scale|k_prob|k_marg|k_grover|k_collapse' -c 24 --csv --log-file $OUT/reductions.csv \
    python scripts/reduction_probe.py > $OUT/reduction_probe.log 2>&1; echo "ncu reductions exit $?"
echo "== parity (everything but the 30-qubit oracle run)"
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -k "not spot_amplitudes_vs_c_oracle" > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
echo "== bench"
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > $OUT/bench.log 2>&1; echo "bench exit $?"; tail -1 $OUT/bench.log | python scripts/bench_brief.py

#!/bin/bash
# One full ncu capture of the tile kernel on the benchmark circuit (28 qubits to keep replay cheap).
TAG=${1:-ncu}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_stage -s ${SKIP:-20} -c ${COUNT:-2} -o $OUT/prof_tile \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --qubits 28 "$@" > $OUT/ncu_full.log 2>&1; echo "ncu full exit $?"; tail -2 $OUT/ncu_full.log | cut -c1-400

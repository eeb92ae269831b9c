#!/bin/bash
TAG=${1:-s7}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for lay in 2x8 2x4 1x8; do
for cfg in 200,4 200,2; do
  IFS=, read c r <<< "$cfg"
  echo "== PROFILE consumers $lay stage-cost $c stage-rounds $r" | tee -a $OUT/prof.log
  QCB_LIB=$PWD/qclojure_b200/lib_prof/libqcb200.so QCB_CONSUMERS=$lay timeout 300 python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --stage-cost $c --stage-rounds $r 2>&1 | grep "tile-prof" | tee -a $OUT/prof.log
done; done

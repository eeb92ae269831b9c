#!/bin/bash
OUT=gpurun_out/${1:-final}; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_gpu.log
timeout 240 python scripts/bench_configs.py > $OUT/configs.json 2> $OUT/configs.err; echo "configs exit $?"; cat $OUT/configs.json; tail -3 $OUT/configs.err

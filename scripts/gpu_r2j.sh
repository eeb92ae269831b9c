#!/bin/bash
TAG=${1:-r2j}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== racecheck probes"
timeout 300 compute-sanitizer --tool racecheck tests/cuda/mbar_racecheck_probe > $OUT/racecheck_probe.log 2>&1; grep -E "Race reported|and (Read|Write)|probe:|SUMMARY" $OUT/racecheck_probe.log | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | cut -c1-200
echo "== noisy probe"
timeout 900 python scripts/noisy_probe.py 2>&1 | tee $OUT/noisy_probe.log

#!/bin/bash
# Round 2, GPU session I (1 GPU): racecheck on the single-tile run with immediate barrier ids, noisy trajectory tree (segment
# form) vs grouping, GPU test-suite, bench.
TAG=${1:-r2i}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== racecheck, single tile"
timeout 900 compute-sanitizer --tool racecheck --print-limit 60 python tests/sanitize_check.py --single-tile > $OUT/racecheck_single_tile.log 2>&1; echo "exit $?"
grep -E "^ok |sanitize_check ok|SUMMARY" $OUT/racecheck_single_tile.log; grep -E "Error: Race|Warning: Race|and (Read|Write) access" $OUT/racecheck_single_tile.log | sed 's/+0x[0-9a-f]*//g; s/\[[0-9]* hazards\]//' | sort | uniq -c | sort -rn | head
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -k "not spot_amplitudes_vs_c_oracle or 28" > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
echo "== other configs"
timeout 600 python scripts/bench_configs.py > $OUT/other_configs.json 2>&1; grep -E "noisy|qaoa" $OUT/other_configs.json
echo "== racecheck, all variants"
timeout 1500 compute-sanitizer --tool racecheck --print-limit 60 python tests/sanitize_check.py > $OUT/sanitizer_racecheck.log 2>&1; grep -E "sanitize_check ok|SUMMARY" $OUT/sanitizer_racecheck.log; grep -E "Error: Race|Warning: Race|and (Read|Write) access" $OUT/sanitizer_racecheck.log | sed 's/+0x[0-9a-f]*//g; s/\[[0-9]* hazards\]//' | sort | uniq -c | sort -rn | head
echo "== bench"
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.log 2>&1; echo "bench exit $?"; tail -1 $OUT/bench.log | python scripts/bench_brief.py
ls -la $OUT

#!/bin/bash
OUT=gpurun_out/r2k; mkdir -p $OUT
timeout 900 compute-sanitizer --tool racecheck --racecheck-report hazard --print-limit 400 python tests/sanitize_check.py --single-tile > $OUT/racecheck_hazards.log 2>&1
grep -c "hazard detected" $OUT/racecheck_hazards.log
grep -E "hazard detected|Thread \(|Access at" $OUT/racecheck_hazards.log | head -80 | cut -c1-220

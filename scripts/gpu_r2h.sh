#!/bin/bash
# Round 2, GPU session H (1 GPU): racecheck probes (mbarrier, named barrier), noisy trajectory tree, marginal / reductions,
# other configs, GPU test-suite (without the 30 q oracle run).
TAG=${1:-r2h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== racecheck probes"
timeout 300 compute-sanitizer --tool racecheck tests/cuda/mbar_racecheck_probe > $OUT/racecheck_probe.log 2>&1; grep -E "Race reported|probe:|SUMMARY" $OUT/racecheck_probe.log | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | cut -c1-200
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -k "not spot_amplitudes_vs_c_oracle or 28" > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
echo "== other configs"
timeout 600 python scripts/bench_configs.py > $OUT/other_configs.json 2>&1; tail -14 $OUT/other_configs.json
QCB_NOISY_TREE=0 timeout 600 python scripts/bench_configs.py > $OUT/other_configs_no_tree.json 2>&1; grep noisy $OUT/other_configs_no_tree.json
echo "== reductions"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'k_marg|k_expect' -c 6 --csv --log-file $OUT/reductions.csv \
    python scripts/reduction_probe.py > $OUT/reduction_probe.log 2>&1; grep -E "k_marginal|k_expect" $OUT/reductions.csv | grep duration | cut -d, -f5,15 | head
echo "== bench quick"
python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --no-hbm-leg --no-other 2>&1 | tail -1 | python scripts/bench_brief.py
ls -la $OUT

#!/bin/bash
# Round 2, GPU session B (2 GPUs): ping-pong loop A/B on GPU 0, then the multi-GPU paths: SPMD parity + group handle parity,
# exchange modes (swap kernel / copy-engine pull / NCCL), 2-GPU bench.
TAG=${1:-r2b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
B="python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --no-hbm-leg"
run() { echo "-- $1" | tee -a $OUT/ab.log; shift; env "$@" 2>&1 | tail -1 | python scripts/bench_brief.py | tee -a $OUT/ab.log; }
echo "== A/B (1 GPU)"
run "k3 ping-pong + realloc r5 (default)" X=1 timeout 300 $B
run "k3 ping-pong, no realloc r5" QCB_LIB=qclojure_b200/lib_var/libqcb200_norealloc.so timeout 300 $B
run "k3 rotate + realloc r5 (session A best)" QCB_LIB=qclojure_b200/lib_var/libqcb200_rotate.so timeout 300 $B
for r in 2 3 4 6 8; do run "default stage-rounds $r" X=1 timeout 300 $B --stage-rounds $r; done
echo "== pytest gpu (quick subset on 1 GPU + multi-GPU tests)"
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 -k "not spot_amplitudes_vs_c_oracle" > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -15 $OUT/pytest_gpu.log
echo "== SPMD check, exchange modes"
for mode in swap ce nccl; do
  QCB_EXCHANGE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > $OUT/check_$mode.log 2>&1; echo "check $mode exit $?"; tail -4 $OUT/check_$mode.log
done
echo "== group check"
timeout 600 python tests/group_check.py 2 > $OUT/group_check.log 2>&1; echo "group exit $?"; tail -6 $OUT/group_check.log
echo "== 2-GPU bench, exchange modes"
for mode in swap ce nccl; do
  echo "-- exchange $mode" | tee -a $OUT/ab.log
  QCB_EXCHANGE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 2 --no-e2e > $OUT/bench2_$mode.log 2>&1
  tail -1 $OUT/bench2_$mode.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('gates/s %.0f ms/step %.1f sweeps %s exchange %s' % (d['value'], d['ms_per_step'], d['sweeps_per_step'], d['exchange']))" | tee -a $OUT/ab.log
done
ls -la $OUT

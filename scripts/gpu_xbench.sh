#!/bin/bash
N=${1:-2}; OUT=gpurun_out/${2:-xbench}; mkdir -p $OUT
run() { echo "== $*" | tee -a $OUT/xbench.log; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 tests/exchange_bench.py 2>&1 | grep "^iter [12]" | tee -a $OUT/xbench.log; }
run QCB_XCHUNK_LOG2=25
run QCB_XCHUNK_LOG2=25 NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32
run QCB_XCHUNK_LOG2=25 NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32 NCCL_BUFFSIZE=16777216
run QCB_XCHUNK_LOG2=25 NCCL_P2P_USE_CUDA_MEMCPY=1
run QCB_XCHUNK_LOG2=27 NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32

#!/bin/bash
# Round 2, GPU session R (1 GPU): Pauli-expectation passes after the select / prefetch rewrite, parity tests that touch them.
TAG=${1:-r2r}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== default"; timeout 300 python scripts/expect_probe.py 30 2>&1 | tee $OUT/expect_default.log
echo "== diagonal loop without prefetch"; QCB_LIB=qclojure_b200/lib_var/libqcb200_olddiag.so timeout 300 python scripts/expect_probe.py 30 2>&1 | tee $OUT/expect_olddiag.log
echo "== parity"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "pauli or qaoa or extraction or golden or hhl" 2>&1 | tail -3 | tee $OUT/pytest_expect.log

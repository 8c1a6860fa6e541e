#!/bin/bash
mkdir -p gpurun_out
echo "== tests =="; timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/tests.txt 2>&1; echo "tests exit $?"; tail -4 gpurun_out/tests.txt
for dbg in 0 1 3 7; do TLB200_TC_DEBUG=$dbg timeout 300 python scripts/prof_time.py 1024 32 2>&1 | tail -1; done
for dbg in 0 3; do TLB200_TC_DEBUG=$dbg timeout 300 python scripts/prof_time.py 768 64 2>&1 | tail -1; done
echo "== ttm_check =="; timeout 600 python scripts/ttm_check.py 512 64 2>&1 | grep -E "auto|chain"
echo "== trace =="; timeout 300 python scripts/tc_trace.py 1024 32 0 2>&1 | tail -8

#!/bin/bash
mkdir -p gpurun_out
echo "== tests =="; timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/tests.txt 2>&1; echo "tests exit $?"; tail -5 gpurun_out/tests.txt
echo "== ttm diag grid=74 =="; TLB200_TC_GRID=74 timeout 300 python scripts/ttm_diag.py 2>&1 | grep -E "^L=" | head -2
for dbg in 0 1 2 3 7; do TLB200_TC_DEBUG=$dbg timeout 300 python scripts/prof_time.py 1024 32 2>&1 | tail -1; done
for dbg in 0 1 2 3; do TLB200_TC_DEBUG=$dbg timeout 300 python scripts/prof_time.py 768 64 2>&1 | tail -1; done
echo "== tc_check C2 =="; timeout 600 python scripts/tc_check.py 1024 32 uniform 2>&1 | tail -3
echo "== ttm_check =="; timeout 600 python scripts/ttm_check.py 512 64 2>&1 | grep -E "auto|chain"
echo "== ncu =="; timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_stream -s 3 -c 1 -o gpurun_out/prof_tc_r1c python scripts/prof_mttkrp.py 1024 32 2 > gpurun_out/ncu_tc.log 2>&1; echo "exit $?"

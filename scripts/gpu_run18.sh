#!/bin/bash
mkdir -p gpurun_out
echo "== tests =="; timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/tests.txt 2>&1; echo "tests exit $?"; tail -3 gpurun_out/tests.txt
timeout 300 python scripts/prof_time.py 1024 32 2>&1 | tail -1
timeout 300 python scripts/prof_time.py 768 64 2>&1 | tail -1
timeout 300 python scripts/prof_time.py 1280 64 2>&1 | tail -1
echo "== ttm_check =="; timeout 600 python scripts/ttm_check.py 512 64 2>&1 | grep -E "auto|chain"

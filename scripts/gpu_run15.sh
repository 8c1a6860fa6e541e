#!/bin/bash
mkdir -p gpurun_out
echo "== tests =="; timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/tests.txt 2>&1; echo "tests exit $?"; tail -3 gpurun_out/tests.txt
timeout 900 python scripts/configs_check.py all 2>&1 | grep -v Warning | tee gpurun_out/configs.txt
echo "== bench short =="; timeout 600 python bench.py --steps 50 --warmup 5 --no-e2e --no-cpu 2>/dev/null | cut -c1-400

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
(timeout 600 python scripts/configs_check.py c3; timeout 900 python scripts/configs_check.py c4) 2>&1 | grep -v Warn | tee gpurun_out/configs_c3_c4.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_stream -s 3 -c 1 -o gpurun_out/prof_ttm_r1 python scripts/prof_ttm.py 512 64 0 > gpurun_out/ncu_ttm.log 2>&1; echo "ncu exit $?"

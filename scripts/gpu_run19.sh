#!/bin/bash
for s in 37 74 18 111 148; do echo "splits $s"; TLB200_TC_SPLITS=$s timeout 300 python scripts/prof_time.py 1024 32 2>&1 | tail -1; done
echo "auto"; timeout 300 python scripts/prof_time.py 1024 32 2>&1 | tail -1
timeout 300 python scripts/prof_time.py 1280 64 2>&1 | tail -1

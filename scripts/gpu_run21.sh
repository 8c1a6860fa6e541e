#!/bin/bash
for d in 0 8; do
TLB200_TC_DEBUG=$d timeout 300 python scripts/prof_time.py 1024 32 2>&1 | tail -1
TLB200_TC_DEBUG=$d timeout 300 python scripts/prof_time.py 1280 64 2>&1 | tail -1
done

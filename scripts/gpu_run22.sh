#!/bin/bash
timeout 300 python scripts/tc_check.py 2>&1 | tail -4
timeout 300 python scripts/prof_time.py 1024 32 2>&1 | tail -1
timeout 300 python scripts/prof_time.py 1280 64 2>&1 | tail -1
timeout 300 python scripts/prof_time.py 768 64 2>&1 | tail -1
TLB200_TC_FLUSH=2 timeout 300 python scripts/prof_time.py 1280 64 2>&1 | tail -1

#!/bin/bash
timeout 300 python scripts/prof_time.py 1280 64 2>&1 | tail -1
timeout 300 python scripts/prof_time.py 768 64 2>&1 | tail -1
timeout 300 python scripts/prof_time.py 1024 32 2>&1 | tail -1
timeout 200 python scripts/tc_trace.py 1280 64 0 2>&1 | tail -6
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3

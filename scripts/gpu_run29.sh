#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -6
TLB200_DIMTREE=0 timeout 300 python scripts/prof_sweep.py 1024 32 2>&1 | grep "graph sweep"
timeout 300 python scripts/prof_sweep.py 1024 32 2>&1 | grep -v "Warn\|warn"
timeout 300 python scripts/prof_sweep.py 768 64 2>&1 | grep -v "Warn\|warn"

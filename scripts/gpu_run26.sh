#!/bin/bash
timeout 60 ./probes/cp_update_probe 32 1024 | tail -1; timeout 60 ./probes/cp_update_probe 64 768 | tail -1; timeout 60 ./probes/cp_update_probe 16 1024 | tail -1; timeout 60 ./probes/cp_update_probe 3 1024 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -5
timeout 300 python scripts/prof_sweep.py 1024 32 2>&1 | grep "graph sweep\|cp_update"
timeout 300 python scripts/prof_sweep.py 768 64 2>&1 | grep "graph sweep\|cp_update"

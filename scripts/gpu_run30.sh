#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -4
timeout 600 python scripts/configs_check.py c4 2>&1 | tail -8
TLB200_DIMTREE=0 timeout 600 python scripts/configs_check.py c4 2>&1 | grep "non_negative_parafac:"
timeout 600 python scripts/configs_check.py c3 2>&1 | tail -5

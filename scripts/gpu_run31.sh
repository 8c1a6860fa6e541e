#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -4
timeout 600 python scripts/configs_check.py c3 2>&1 | tail -6

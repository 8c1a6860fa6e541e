#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/tests23.txt 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/tests23.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2

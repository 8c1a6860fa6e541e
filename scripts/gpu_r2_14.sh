#!/bin/bash
# visit 14: fp16-split engine bring-up
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "fp16_engine" > gpurun_out/tests14.txt 2>&1; echo "tests rc=$?"
tail -30 gpurun_out/tests14.txt

#!/bin/bash
# round-2 second GPU visit: the whole GPU suite (new recon / masked / HOOI / solve-hook tests), then a short bench
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/tests2.txt 2>&1; echo "tests rc=$?" >> gpurun_out/tests2.txt
timeout 900 python bench.py --no-e2e --no-cpu > gpurun_out/bench2_c5_n1.json 2> gpurun_out/bench2_c5_n1.err; echo "bench rc=$?"
grep -v "^$" gpurun_out/tests2.txt | tail -n 40; tail -n 5 gpurun_out/bench2_c5_n1.err

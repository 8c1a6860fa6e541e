#!/bin/bash
mkdir -p gpurun_out
./probes/dmma_rate > gpurun_out/dmma_rate.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_c5slab.csv python bench.py --workload c5slab --steps 3 --warmup 3 --no-e2e --no-cpu --no-c2 --no-c3 --no-fp64 --no-refdriver --no-sustained > gpurun_out/prof_c5slab.log 2>&1; echo "ncu rc=$?"
python scripts/launch_summary.py gpurun_out/launches_c5slab.csv > gpurun_out/launches_c5slab_summary.txt 2>&1
timeout 600 python bench.py --workload c5slab --no-e2e --no-cpu --no-c2 --no-fp64 --no-refdriver > gpurun_out/bench_c5slab.json 2> gpurun_out/bench_c5slab.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_c3c.csv python scripts/prof_c3.py > gpurun_out/prof_c3c.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_c3c.csv > gpurun_out/launches_c3c_summary.txt 2>&1
cat gpurun_out/dmma_rate.txt; cat gpurun_out/launches_c5slab_summary.txt | head -20; head -14 gpurun_out/launches_c3c_summary.txt
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_c5slab.json').read().strip().splitlines()[-1])
print('c5slab', d['value'], d['ms_per_step'], d['launches_per_sweep'], d['roofline']['per_mode_gbs'], d['ttm_pass'])
c3=d['c3']; print('c3', c3['value'], c3['ms_per_step'], c3['launches_per_sweep'], c3.get('parity_vs_reference_driver'))
P

#!/bin/bash
# visit 15: fp16-split engine A/B on the same box (C5 headline + c5slab)
mkdir -p gpurun_out
for hf in 0 1; do
  TLB200_DISABLE_HF=$hf timeout 900 python bench.py --workload c5 --steps 20 --warmup 5 --no-e2e --no-cpu --no-fp64 --no-refdriver --no-c3 --no-c2 > gpurun_out/bench15_c5_dis$hf.json 2> gpurun_out/bench15_c5_dis$hf.err; echo "bench c5 disable_hf=$hf rc=$?"
  TLB200_DISABLE_HF=$hf timeout 600 python bench.py --workload c5slab --steps 40 --warmup 5 --no-e2e --no-cpu --no-fp64 --no-refdriver --no-c3 --no-c2 > gpurun_out/bench15_slab_dis$hf.json 2> gpurun_out/bench15_slab_dis$hf.err; echo "bench slab disable_hf=$hf rc=$?"
done
python - <<'P'
import json
for n in ('c5_dis0','c5_dis1','slab_dis0','slab_dis1'):
    try:
        d=json.loads(open(f'gpurun_out/bench15_{n}.json').read().strip().splitlines()[-1])
        print(n, round(d['value'],2), round(d['ms_per_step'],3), d['roofline'], d.get('clocks'), d.get('parity'), d.get('sustained'))
    except Exception as e:
        print(n, 'failed', e)
P

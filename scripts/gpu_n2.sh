#!/bin/bash
# 2-GPU check of the sharded driver (peer-memory all-reduce, parity with the single-GPU trajectory), then the bench command
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1"
mkdir -p gpurun_out
echo "== dist check (2 GPUs) =="; timeout 300 $TR --master-port 29511 scripts/dist_check.py > gpurun_out/dist_check_n${NG:-2}.txt 2>&1; echo "exit $?"; grep -E "^rank|Error|error|Traceback" gpurun_out/dist_check_n${NG:-2}.txt | head -20
echo "== bench 2 GPUs (driver command) =="; timeout 600 $TR --master-port 29512 bench.py --gpus ${NG:-2} --steps 20 --warmup 5 > gpurun_out/bench_n${NG:-2}.json 2> gpurun_out/bench_n${NG:-2}.err; echo "exit $?"; cut -c1-400 gpurun_out/bench_n${NG:-2}.json; grep -iE "error|Traceback" gpurun_out/bench_n${NG:-2}.err | head -5
python - <<'P'
import json
try:
    import os
    d=json.loads(open('gpurun_out/bench_n%s.json' % os.environ.get('NG','2')).read().strip().splitlines()[-1])
    print('c5', d['value'], d['ms_per_step'], d['config']['collective'], d['parity'], d['launches_per_sweep'])
    print('c2', d['c2']['value'], d['c2']['ms_per_step'])
    print('e2e', d['e2e'])
except Exception as e: print('parse failed', e)
P

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "cp_to_tensor or impute or masked" > gpurun_out/tests34.txt 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/tests34.txt
python - <<'P'
import torch, time
x = torch.empty(1<<30, dtype=torch.float32, device='cuda')
for name, fn in (('fill_', lambda: x.fill_(1.5)), ('zero_', lambda: x.zero_())):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): fn()
    b.record(); torch.cuda.synchronize()
    print(name, 'write-only GB/s', x.numel()*4/(a.elapsed_time(b)/10*1e-3)/1e9)
P

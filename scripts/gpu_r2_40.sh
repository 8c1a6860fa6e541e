#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "cp_to_tensor" 2>&1 | tail -2
for w in 2 4; do
TLB200_RECON_LINES=$w python - <<'P'
import os, torch, tensorly_b200 as tb
shape, R = (512, 1024, 1024), 32
g = torch.Generator(device="cuda").manual_seed(9)
fs = [torch.rand((s, R), generator=g, device="cuda") for s in shape]
w = torch.ones(R, device="cuda"); out = torch.empty(shape, device="cuda")
for _ in range(3): tb.cp_to_tensor((w, fs), out=out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): tb.cp_to_tensor((w, fs), out=out)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
print("lines", os.environ["TLB200_RECON_LINES"], "cp_to_tensor %.3f ms  %.0f GB/s" % (ms, out.numel() * 4 / ms / 1e6))
P
done

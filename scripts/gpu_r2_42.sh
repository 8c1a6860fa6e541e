#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "cp_to_tensor or impute or masked" 2>&1 | tail -2
python - <<'P'
import torch, tensorly_b200 as tb
for shape, R in [((128, 128, 128, 128), 32), ((128, 128, 128, 128), 64), ((512, 1024, 1024), 32), ((512, 1024, 1024), 64)]:
    g = torch.Generator(device="cuda").manual_seed(9)
    fs = [torch.rand((s, R), generator=g, device="cuda") for s in shape]
    w = torch.ones(R, device="cuda")
    x = torch.rand(shape, generator=g, device="cuda"); mask = (torch.rand(shape, generator=g, device="cuda") > 0.1).float()
    out = torch.empty(shape, device="cuda")
    res = []
    for name, fn in (("recon", lambda: tb.cp_to_tensor((w, fs), out=out)), ("impute", lambda: tb.cp_impute(x, mask, (w, fs), out=x))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): fn()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        by = out.numel() * (4 if name == "recon" else 12)
        res.append(f"{name} {ms:.3f} ms {by / ms / 1e6:.0f} GB/s")
    print(shape, R, tb.last_kernel_path(), " | ".join(res))
P

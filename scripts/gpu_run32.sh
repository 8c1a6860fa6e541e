#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -4
timeout 300 python - <<'PY'
import torch, time, sys
sys.path.insert(0, '.')
import tensorly_b200 as tb
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.rand(768, 768, 768, generator=g, device="cuda")
for R in (64, 100, 128):
    fs = [torch.rand(768, R, generator=g, device="cuda") for _ in range(3)]
    for path in ("auto", "simt"):
        tb.set_kernel_path(path)
        for _ in range(2): tb.unfolding_dot_khatri_rao(x, (None, fs), 1)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3): tb.unfolding_dot_khatri_rao(x, (None, fs), 1)
        torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 3 * 1e3
        print(f"MTTKRP 768^3 R={R} mode 1 {path:5s} ({tb.last_kernel_path()}): {ms:.3f} ms  {x.numel()*4/ms/1e6:.0f} GB/s")
    m = torch.rand(R, 768, generator=g, device="cuda")
    for path in ("auto", "simt"):
        tb.set_kernel_path(path)
        for _ in range(2): tb.mode_dot(x, m, 1)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3): tb.mode_dot(x, m, 1)
        torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 3 * 1e3
        print(f"mode_dot 768^3 x ({R},768) mode 1 {path:5s} ({tb.last_kernel_path()}): {ms:.3f} ms")
PY

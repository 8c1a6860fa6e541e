"""Dimension-tree tensor pass T = X x_last F^T (rows = 1M, contraction 2048, 64 columns): 3xTF32 vs fp16-split engine."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tensorly_b200 as tb
from bench import ClockSampler

shape = (512, 2048, 2048)
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.rand(shape, generator=g, device="cuda")
f = torch.rand(2048, 64, generator=g, device="cuda")
cs = ClockSampler(0); cs.start(); time.sleep(0.3)
ref = None
for use_hint in (False, True, False, True):
    hint = tb.RangeHint(x) if use_hint else None
    for _ in range(3):
        t = tb.mode_dot(x, f, 2, transpose=True)
    torch.cuda.synchronize()
    path = tb.last_kernel_path()
    if ref is None:
        ref = t.clone()
    dev = float((t - ref).norm() / ref.norm())
    # burst: 10 calls
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        tb.mode_dot(x, f, 2, transpose=True)
    e1.record(); torch.cuda.synchronize()
    burst = e0.elapsed_time(e1) / 10
    m0 = cs.mark(); n = 0; t0 = time.time(); e0.record()
    while time.time() - t0 < 2.0:
        for _ in range(20):
            tb.mode_dot(x, f, 2, transpose=True)
        n += 20; torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    s = cs.summary(m0, cs.mark())
    by = x.numel() * 4 + t.numel() * 4
    print(f"hint={use_hint} [{path}] burst {by / burst / 1e6:6.0f} GB/s ({burst:.3f} ms)  sustained {by / ms / 1e6:6.0f} GB/s  sm {s['sm_mhz']} MHz "
          f"{s['power_w_max']} W  dev vs first {dev:.2e}", flush=True)
    if hint: hint.close()
cs.stop()

"""Per-kernel device times of one CP-ALS sweep (torch.profiler / CUPTI, eager launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tensorly_b200 as tb
from tensorly_b200.cp_als import CPALS
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
R = int(sys.argv[2]) if len(sys.argv) > 2 else 32
n0 = int(sys.argv[3]) if len(sys.argv) > 3 else n          # a mode-0 slab (what one rank of a sharded run holds)
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.rand(n0, n, n, generator=g, device="cuda")
fs = [torch.rand(s, R, generator=g, device="cuda") for s in (n0, n, n)]
w = torch.ones(R, device="cuda")
st = CPALS(x, w, fs)
for _ in range(3): st.sweep_eager()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5): st.sweep_eager()
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", None) or getattr(e, "cuda_time_total", 0)
    if t: rows.append((t / 5, e.count / 5, e.key[:90]))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"shape=({n0},{n},{n}) R={R}: kernel time per sweep {tot/1e3:.3f} ms")
for t, c, k in rows: print(f"  {t:9.1f} us/sweep  x{c:4.1f}  {t/c:8.1f} us each  {k}")
# graph replay time
for _ in range(3): st.sweep()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(20): st.sweep()
e1.record(); torch.cuda.synchronize()
print(f"graph sweep {e0.elapsed_time(e1)/20:.3f} ms")

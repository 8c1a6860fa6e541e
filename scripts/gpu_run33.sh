#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
timeout 600 python - <<'PY'
import torch, time, sys
sys.path.insert(0, '.')
import tensorly_b200 as tb
g = torch.Generator(device="cuda").manual_seed(0)
def timeit(f, n=3):
    for _ in range(2): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
R = 64
x = torch.rand(256, 256, 256, 256, generator=g, device="cuda")
fs = [torch.rand(256, R, generator=g, device="cuda") for _ in range(4)]
ms = timeit(lambda: tb.mode_dot(x, fs[3], 3, transpose=True))
print(f"C4 TTM pass: {ms:.3f} ms  {(x.numel()*4 + x.numel()//256*R*4)/ms/1e6:.0f} GB/s")
t = tb.mode_dot(x, fs[3], 3, transpose=True)
for mode in range(3):
    ms = timeit(lambda: tb.mttkrp_from_ttm(t, (None, fs), mode))
    print(f"C4 from_ttm mode {mode}: {ms:.3f} ms  {t.numel()*4/ms/1e6:.0f} GB/s")
for mode in range(4):
    ms = timeit(lambda: tb.unfolding_dot_khatri_rao(x, (None, fs), mode))
    print(f"C4 MTTKRP mode {mode}: {ms:.3f} ms {x.numel()*4/ms/1e6:.0f} GB/s")
del t
def run(n): return tb.non_negative_parafac(x, R, n_iter_max=n, init=(None, fs), tol=0)
run(3); torch.cuda.synchronize(); t0 = time.perf_counter(); run(3); torch.cuda.synchronize(); ta = time.perf_counter() - t0
t0 = time.perf_counter(); run(13); torch.cuda.synchronize(); tb_ = time.perf_counter() - t0
print(f"C4 NN-CP: {10/(tb_-ta):.1f} sweeps/s  (ta {ta*1e3:.1f} ms, tb {tb_*1e3:.1f} ms)")
PY

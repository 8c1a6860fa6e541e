import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tensorly_b200 as tb
tl = tb.import_tensorly(); tl.set_backend("pytorch"); tb.use()
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.rand(512, 512, 512, generator=g, device="cuda")
us = [torch.randn(512, 64, generator=g, device="cuda").t().contiguous().t() for _ in range(3)]
def timeit(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) / n * 1e3
for name, fn in (("tb.multi_mode_dot skip=0", lambda: tb.multi_mode_dot(x, us, skip=0, transpose=True)),
                 ("tl.tenalg.multi_mode_dot skip=0", lambda: tl.tenalg.multi_mode_dot(x, us, skip=0, transpose=True)),
                 ("tl.tenalg.multi_mode_dot full", lambda: tl.tenalg.multi_mode_dot(x, us, transpose=True)),
                 ("4 chains list", lambda: [tl.tenalg.multi_mode_dot(x, us, skip=k, transpose=True) for k in range(3)] + [tl.tenalg.multi_mode_dot(x, us, transpose=True)])):
    ev, wall = timeit(fn)
    print(f"{name:36s} events {ev:8.3f} ms  wall {wall:8.3f} ms  path {tb.last_kernel_path()}")
us2 = [torch.randn(512, 64, generator=g, device="cuda") for _ in range(3)]   # row-major (n, I)
ev, wall = timeit(lambda: tb.multi_mode_dot(x, us2, skip=0, transpose=True))
print(f"row-major factors skip=0             events {ev:8.3f} ms  wall {wall:8.3f} ms")

"""Where does the power go?  MTTKRP (2048 x 2048 x 256, R = 64, each mode) timed for ~2 s per variant with parts of
the engine switched off (TLB200_TC_DEBUG: 2 = no MMAs, 4 = no convert math / TMEM stores), nvidia-smi clocks and power
sampled meanwhile.  Run once per (TLB200_DISABLE_HF, TLB200_TC_DEBUG) setting — both are read once per process."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import tensorly_b200 as tb
from bench import ClockSampler

rank = int(os.environ.get("RANK_R", "64"))
shape = (2048, 2048, 256)
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.rand(shape, generator=g, device="cuda")
fs = [torch.rand(s, rank, generator=g, device="cuda") for s in shape]
hint = tb.RangeHint(x) if os.environ.get("TLB200_DISABLE_HF", "0") == "0" else None
cs = ClockSampler(0)
cs.start()
time.sleep(0.3)
out = {}
for mode in range(3):
    for _ in range(5):
        tb.unfolding_dot_khatri_rao(x, (None, fs), mode)
    torch.cuda.synchronize()
    m0 = cs.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    t0 = time.time()
    e0.record()
    while time.time() - t0 < 2.0:
        for _ in range(50):
            tb.unfolding_dot_khatri_rao(x, (None, fs), mode)
        n += 50
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    s = cs.summary(m0, cs.mark())
    print(f"HFoff={os.environ.get('TLB200_DISABLE_HF', '0')} debug={os.environ.get('TLB200_TC_DEBUG', '0')} R={rank} mode {mode} "
          f"[{tb.last_kernel_path()}]: {x.numel() * 4 / ms / 1e6:7.0f} GB/s  {ms * 1e3:7.1f} us  sm {s['sm_mhz']} MHz  "
          f"mem {s.get('mem_mhz')} MHz  {s.get('temp_c_max')} C  power max {s['power_w_max']} W  {s['reasons']}", flush=True)
# the ceiling under the same power cap: a read-only pass over the same tensor (max |x|), sustained
for name, fn in (("tensor_absmax (read-only pass)", lambda: tb.tensor_absmax(x)), ("sumsq (read-only pass)", lambda: tb.sumsq(x))):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    m0 = cs.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    t0 = time.time()
    e0.record()
    while time.time() - t0 < 2.0:
        for _ in range(50):
            fn()
        n += 50
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    s = cs.summary(m0, cs.mark())
    print(f"ceiling {name}: {x.numel() * 4 / ms / 1e6:7.0f} GB/s  {ms * 1e3:7.1f} us  sm {s['sm_mhz']} MHz  mem {s.get('mem_mhz')} MHz  {s.get('temp_c_max')} C  power max {s['power_w_max']} W  "
          f"{s['reasons']}", flush=True)
cs.stop()

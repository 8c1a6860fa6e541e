"""CTA 0's pipeline timeline of the tcgen05 engine (in-kernel clock64 stamps): python scripts/tc_trace.py [n] [rank] [mode]
Roles: 0 X TMA producer, 1 / 3 the two MMA issuers (alternate tiles), 2 convert warp 2 (set 0: every other tile),
4 epilogue warp 12 (per accumulation group)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import tensorly_b200 as tb
from tensorly_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
R = int(sys.argv[2]) if len(sys.argv) > 2 else 32
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
lib = _lib.load()
lib.tlb200_debug_set_trace.argtypes = [ctypes.c_void_p]
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.rand(n, n, n, generator=g, device="cuda")
fs = [torch.rand(n, R, generator=g, device="cuda") for _ in range(3)]
w = torch.ones(R, device="cuda")
for _ in range(2): tb.unfolding_dot_khatri_rao(x, (w, fs), mode)
buf = torch.zeros(5 * 256 * 8, dtype=torch.int64, device="cuda")
lib.tlb200_debug_set_trace(buf.data_ptr())
tb.unfolding_dot_khatri_rao(x, (w, fs), mode)
torch.cuda.synchronize()
lib.tlb200_debug_set_trace(None)
t = buf.cpu().numpy().reshape(5, 256, 8).astype(np.float64)
def valid(role, ev): return int((t[role, :, ev] > 0).sum())
print(f"n={n} R={R} mode={mode}: stamped iterations per role:", {r: valid(r, 0) for r in range(5)})
lo, hi = 20, 100
tm = t[0, lo:hi]
print("TMA producer : tile period %.0f clk, wait for a free stage %.0f, issue %.0f" % (np.diff(tm[:, 0]).mean(), (tm[1:, 0] - tm[:-1, 1]).mean(), (tm[:, 1] - tm[:, 0]).mean()))
for role, name in ((1, "MMA issuer 0"), (3, "MMA issuer 1")):
    m = t[role, lo // 2:hi // 2]
    print("%s : own-tile period %.0f clk (= 2 tiles), wait operands+turn %.0f, issue+commit %.0f, skip the other's tile %.0f, bookkeeping + accumulator-free wait %.0f (max %.0f)" % (
        name, np.diff(m[:, 0]).mean(), (m[:, 2] - m[:, 0]).mean(), (m[:, 3] - m[:, 2]).mean(), (m[1:, 4] - m[:-1, 3]).mean(),
        (m[:, 0] - m[:, 4]).mean(), (m[:, 0] - m[:, 4]).max()))
c = t[2, lo // 2:hi // 2]
print("convert set 0: own-tile period %.0f clk (= 2 tiles), wait x_full %.0f, smem->regs+release %.0f, wait a_empty %.0f, split+st %.0f, wait::st+publish %.0f" % (
    np.diff(c[:, 0]).mean(), (c[:, 1] - c[:, 0]).mean(), (c[:, 2] - c[:, 1]).mean(), (c[:, 4] - c[:, 3]).mean(), (c[:, 6] - c[:, 4]).mean() - 0, 0))
e = t[4, 4:valid(4, 0) - 1] if valid(4, 0) > 8 else t[4, 1:valid(4, 0)]
print("epilogue     : group period %.0f clk, wait d_full %.0f, drain %.0f" % (np.diff(e[:, 0]).mean(), (e[:, 1] - e[:, 0]).mean(), (e[:, 2] - e[:, 1]).mean()))

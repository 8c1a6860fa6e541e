"""Dump CTA 0's pipeline timeline of the tcgen05 engine: python scripts/tc_trace.py [n] [rank] [mode]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
t = buf.cpu().numpy().reshape(5, 256, 8)
t0 = t[t > 0].min()
def rel(v): return int(v - t0) if v > 0 else -1
names = {0: ["x_empty ok", "tma issued"], 1: ["mma top", "a_full ok", "b_full ok", "issued+commit"],
         2: ["cv top", "x_full ok", "x released", "published", "a_empty0 ok", "a_empty1 ok", "sttm issued"],
         3: ["kr staged", "b_empty ok", "b_full arr"], 4: ["epi top", "d_full ok", "drained"]}
lo, hi = 96, 104
print("TMA (per chunk):")
for i in range(lo // 2, hi // 2): print("  ", i, [rel(v) for v in t[0, i, :2]], "dt", int(t[0, i, 0] - t[0, i - 1, 0]))
print("convert warp 2 (per chunk):", names[2])
for i in range(lo // 2, hi // 2): print("  ", i, [rel(v) for v in t[2, i, :7]], "dt", int(t[2, i, 0] - t[2, i - 1, 0]))
print("MMA (per unit):", names[1])
for i in range(lo, hi): print("  ", i, [rel(v) for v in t[1, i, :4]], "dt", int(t[1, i, 0] - t[1, i - 1, 0]))
print("KR warp 6 (every 4th unit):", names[3])
for i in range(lo // 4, hi // 4 + 1): print("  ", i, [rel(v) for v in t[3, i, :3]], "dt", int(t[3, i, 0] - t[3, i - 1, 0]))
print("epilogue warp 10 (per group):", names[4])
for i in range(lo // 8, hi // 8 + 1): print("  ", i, [rel(v) for v in t[4, i, :3]], "dt", int(t[4, i, 0] - t[4, i - 1, 0]))
import numpy as np
d = np.diff(t[2, 20:200, 0]); print("convert chunk period: mean %.0f min %d max %d" % (d.mean(), d.min(), d.max()))
d = np.diff(t[1, 40:250, 0]); print("mma unit period: mean %.0f" % d.mean())
print("mma: wait a_full mean %.0f  wait b_full mean %.0f  issue mean %.0f" % ((t[1, 40:250, 1] - t[1, 40:250, 0]).mean(), (t[1, 40:250, 2] - t[1, 40:250, 1]).mean(), (t[1, 40:250, 3] - t[1, 40:250, 2]).mean()))
c = t[2, 20:200]
print("convert: wait x_full %.0f  load+release %.0f  publish %.0f  a_empty0 %.0f  a_empty1(+sttm0) %.0f  sttm1 %.0f" % ((c[:, 1] - c[:, 0]).mean(), (c[:, 2] - c[:, 1]).mean(), (c[:, 3] - c[:, 2]).mean(), (c[:, 4] - c[:, 3]).mean(), (c[:, 5] - c[:, 4]).mean(), (c[:, 6] - c[:, 5]).mean()))
k = t[3, 10:60]
print("KR warp: wait b_empty %.0f  compute+arrive %.0f  period %.0f" % ((k[:, 1] - k[:, 0]).mean(), (k[:, 2] - k[:, 1]).mean(), np.diff(k[:, 0]).mean()))
tm = t[0, 20:120]
print("TMA: wait x_empty %.0f period %.0f" % ((tm[1:, 0] - tm[:-1, 1]).mean(), np.diff(tm[:, 0]).mean()))

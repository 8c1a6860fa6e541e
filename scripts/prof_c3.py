"""A few HOOI sweeps at C3 (512^3, ranks 64) for an ncu launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c3.csv python scripts/prof_c3.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tensorly_b200 as tb
x = torch.rand((512, 512, 512), device="cuda")
ranks = [64, 64, 64]
fs = tb.tucker_hooi._svd_init(tb.tucker_hooi.CudaOps, x, ranks, [0, 1, 2])
st = tb.HOOI(x, ranks, [0, 1, 2], fs)
for _ in range(int(os.environ.get("SWEEPS", "2"))):
    st.sweep()
torch.cuda.synchronize()
print("err", float(st.err[0]))

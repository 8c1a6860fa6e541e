"""torchrun check of the mode-sharded CP-ALS on real GPUs: sharded trajectory == single-GPU trajectory."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import tensorly_b200 as tb
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
shape, R = (384, 320, 256), 32
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.rand(shape, generator=g, device="cuda")
fs = [torch.rand(s, R, generator=g, device="cuda") for s in shape]
lo, hi = tb.shard_bounds(shape[0], world, rank)
cp, errs = tb.parafac(x[lo:hi].contiguous(), R, n_iter_max=6, init=(None, fs), tol=0, return_errors=True, sharded=True, shard_mode=0)
if rank == 0:
    dist_errs = errs
torch.cuda.synchronize()
# single-GPU reference on every rank (no group): temporarily pretend world size 1 by using ops directly
from tensorly_b200.cp_als import CPALS, _Comm
class NoComm(_Comm):
    def __init__(self): self.active = False; self.world = 1; self.rank = 0
st = CPALS(x, torch.ones(R, device="cuda"), fs, comm=NoComm())
ref = []
for _ in range(6):
    st.sweep_eager(True); ref.append(float(st.err[0]))
dev = max(abs(a - b) / b for a, b in zip(errs, ref))
fdev = max(float(torch.linalg.norm(a - b) / torch.linalg.norm(b)) for a, b in zip(cp[1], st.factors))
print(f"rank {rank}/{world}: sharded vs single-GPU rel-error deviation {dev:.2e}, factor deviation {fdev:.2e}, errs {errs[:3]}")
assert dev < 1e-4 and fdev < 1e-2
dist.destroy_process_group()

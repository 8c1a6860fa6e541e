"""torchrun check of the sharded CP-ALS on real GPUs:
  1. the peer-memory all-reduce (csrc/comm.cu) against NCCL's, many sizes, many epochs, and replayed from a CUDA graph;
  2. sharded trajectory == single-GPU trajectory (eager and captured sweeps);
  3. sharded masked ALS == single-GPU masked ALS."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import tensorly_b200 as tb
from tensorly_b200.cp_als import CPALS, _Comm

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
comm = _Comm(None, sharded=True)
probe = torch.ones(4, device=dev)
comm.all_reduce(probe)
assert torch.equal(probe, torch.full((4,), float(world), device=dev)), probe
print(f"rank {rank}: collective = {comm.kind}", flush=True)

# 1. all-reduce vs NCCL
g = torch.Generator(device=dev).manual_seed(100 + rank)
worst = 0.0
for it in range(40):
    for n, dt in ((1, torch.float32), (3, torch.float64), (4096, torch.float32), (2048 * 64 + 64 * 64, torch.float32),
                  (100003, torch.float64), (7, torch.float32)):
        a = torch.randn(n, generator=g, device=dev, dtype=dt)
        ref = a.clone()
        dist.all_reduce(ref)
        got = comm.all_reduce(a.clone())
        worst = max(worst, float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30)))
        # same bits on every rank
        chk = got.double().sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert float(lo) == float(hi), (n, dt)
assert worst < 1e-5, worst
if comm.graph_safe:
    buf = torch.zeros(70000, device=dev)
    src = torch.randn(70000, generator=g, device=dev)
    torch.cuda.synchronize(); dist.barrier()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        buf.copy_(src)
        comm.all_reduce(buf)
        comm.all_reduce(buf)
    ref = src.clone(); dist.all_reduce(ref); ref *= world
    for _ in range(25):
        gr.replay()
    torch.cuda.synchronize()
    assert float((buf - ref).abs().max() / ref.abs().max()) < 1e-5
    del gr
print(f"rank {rank}: all-reduce ok (max rel dev vs NCCL {worst:.1e})", flush=True)

# 2. sharded vs single-GPU trajectories
shape, R = (384, 320, 256), 32
g = torch.Generator(device=dev).manual_seed(0)
x = torch.rand(shape, generator=g, device=dev)
fs = [torch.rand(s, R, generator=g, device=dev) for s in shape]
lo, hi = tb.shard_bounds(shape[0], world, rank)
solo = _Comm(None, sharded=False)
st = CPALS(x, torch.ones(R, device=dev), fs, comm=solo)
ref = []
for _ in range(6):
    st.sweep(True); ref.append(float(st.err[0]))
for use_graph in (False, True):
    cp, errs = tb.parafac(x[lo:hi].contiguous(), R, n_iter_max=6, init=(None, fs), tol=0, return_errors=True, sharded=True,
                          shard_mode=0, use_graph=use_graph)
    d = max(abs(a - b) / b for a, b in zip(errs, ref))
    fdev = max(float(torch.linalg.norm(a - b) / torch.linalg.norm(b)) for a, b in zip(cp[1], st.factors))
    print(f"rank {rank}/{world}: graph={use_graph} sharded vs single-GPU rel-error deviation {d:.2e}, factor deviation {fdev:.2e}", flush=True)
    assert d < 1e-4 and fdev < 1e-2

# 3. masked
mask = (torch.rand(shape, generator=g, device=dev) > 0.2).float()
xm = x * mask
_, e1 = tb.parafac(xm, R, n_iter_max=4, init=(None, fs), tol=0, return_errors=True, mask=mask)
_, e2 = tb.parafac(xm[lo:hi].contiguous(), R, n_iter_max=4, init=(None, fs), tol=0, return_errors=True, mask=mask[lo:hi].contiguous(),
                   sharded=True, shard_mode=0)
d = max(abs(a - b) / b for a, b in zip(e2, e1))
print(f"rank {rank}: masked sharded vs single deviation {d:.2e}", flush=True)
assert d < 1e-4
torch.cuda.synchronize(); dist.barrier()
comm.close()
dist.destroy_process_group()
print(f"rank {rank}: done", flush=True)

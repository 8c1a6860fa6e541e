"""world_size-2 gloo tests (CPU) of the mode-sharded CP-ALS driver.

The compute ops are swapped for oracle-backed CPU stand-ins (tests/oracle_ops.py), so what
is under test is the product's host logic: slab partition, which MTTKRP/Gram partials are
all-reduced, the error assembly, the final gather — checked against the single-process
reference algorithm on the same tensor and the same initial factors."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, shape, cp_rank, shard_mode, update, iters, ret, masked=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import tensorly_b200 as tb
        from oracle import oracle as O
        from oracle_ops import OracleOps

        calls = [0]
        real_all_reduce = dist.all_reduce

        def counting_all_reduce(*a, **k):
            calls[0] += 1
            return real_all_reduce(*a, **k)
        dist.all_reduce = counting_all_reduce
        x = O.random_tensor(shape, 0)
        w, fs = O.random_cp_factors(shape, cp_rank, 1)
        lo, hi = tb.shard_bounds(shape[shard_mode], world, rank)
        sl = [slice(None)] * len(shape)
        sl[shard_mode] = slice(lo, hi)
        x_local = torch.from_numpy(np.ascontiguousarray(x[tuple(sl)]))
        init = (None, [torch.from_numpy(f.copy()) for f in fs])
        fn = tb.parafac if update == "ls" else tb.non_negative_parafac
        kw = dict(n_iter_max=iters, init=init, return_errors=True, sharded=True, shard_mode=shard_mode, ops=OracleOps, use_graph=False)
        kw["tol"] = 0 if update == "ls" else 1e-30
        if masked:
            mask = (np.random.RandomState(5).random_sample(shape) > 0.25).astype(x.dtype)
            x = x * mask
            x_local = torch.from_numpy(np.ascontiguousarray(x[tuple(sl)]))
            kw["mask"] = torch.from_numpy(np.ascontiguousarray(mask[tuple(sl)]))
        cp, errs = fn(x_local, cp_rank, **kw)
        if masked:
            (_, ref_f), ref_e, _ = O.parafac_masked(x, mask, (w, fs), n_iter_max=iters)
        elif update == "ls":
            (_, ref_f), ref_e = O.parafac(x, (w, fs), n_iter_max=iters)
        else:
            (_, ref_f), ref_e = O.non_negative_parafac(x, (w, fs), n_iter_max=iters)
        err_dev = float(np.max(np.abs(np.array(errs) - np.array(ref_e)) / np.array(ref_e)))
        fac_dev = max(float(np.linalg.norm(a.numpy() - b) / np.linalg.norm(b)) for a, b in zip(cp[1], ref_f))
        shapes_ok = all(tuple(a.shape) == b.shape for a, b in zip(cp[1], ref_f))
        ret[rank] = (err_dev, fac_dev, shapes_ok, len(errs), calls[0])
    finally:
        dist.destroy_process_group()


def _run(shape, cp_rank, shard_mode, update="ls", iters=4, world=2, all_reduces=None, masked=False):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), shape, cp_rank, shard_mode, update, iters, ret, masked), nprocs=world,
             join=True)
    assert len(ret) == world
    for r in range(world):
        err_dev, fac_dev, shapes_ok, n, n_all_reduce = ret[r]
        if all_reduces is not None:
            assert n_all_reduce == all_reduces, (r, n_all_reduce)
        assert shapes_ok and n == iters
        assert err_dev <= 1e-9, (r, err_dev)
        assert fac_dev <= 1e-7, (r, fac_dev)


@pytest.mark.timeout(300)
def test_sharded_parafac_mode0_matches_single_process():
    # collectives: ||X||^2 + the initial Gram of the sharded mode, then per sweep ONE packed all-reduce (mode-0 Gram
    # partial + mode-1 MTTKRP partial) and one for the mode-2 MTTKRP: 2 + 4 * 2
    _run((12, 9, 10), 3, shard_mode=0, all_reduces=10)


@pytest.mark.timeout(300)
def test_sharded_parafac_uneven_slabs_and_middle_mode():
    _run((7, 11, 6), 2, shard_mode=0)        # 4 + 3 rows
    _run((6, 9, 5), 2, shard_mode=1)


@pytest.mark.timeout(300)
def test_sharded_parafac_last_mode_and_four_way():
    # last mode sharded: nothing follows it inside a sweep, so its Gram travels alone, and iprod is a partial sum:
    # 2 + 4 * (M0, M1, Gram, iprod)
    _run((5, 6, 9), 2, shard_mode=2, all_reduces=18)
    _run((4, 5, 6, 7), 3, shard_mode=0)


@pytest.mark.timeout(300)
def test_sharded_non_negative_parafac():
    _run((8, 7, 6), 3, shard_mode=0, update="mu", iters=5)


@pytest.mark.timeout(300)
def test_sharded_masked_parafac():
    """Missing values + sharding: every rank imputes its own slab, the two norms are all-reduced."""
    _run((10, 8, 9), 3, shard_mode=0, masked=True)


def test_masked_parafac_host_logic_vs_reference_golden():
    """Own driver with `mask=` (oracle-backed ops, fp64) against the reference's masked parafac trajectories; the
    caller's tensor is not modified; plain sharded=False inside a process without a process group stays local."""
    import tensorly_b200 as tb
    from oracle_ops import OracleOps
    from conftest import Golden
    g = Golden("round2")
    for tag in ("mask64", "mask4way"):
        x, mask = g[f"{tag}/x"], g[f"{tag}/mask"]
        rank, iters = int(g[f"{tag}/rank"]), int(g[f"{tag}/iters"])
        xt = torch.from_numpy(x.copy())
        init = (None, [torch.from_numpy(f.copy()) for f in g.arrays(tag, "init")])
        cp, errs = tb.parafac(xt, rank, n_iter_max=iters, init=init, tol=0, return_errors=True,
                              mask=torch.from_numpy(mask.copy()), ops=OracleOps, use_graph=False)
        ref = g[f"{tag}/errors"]
        assert np.max(np.abs(np.array(errs) - ref) / ref) <= 1e-10
        assert np.array_equal(xt.numpy(), x)
        for a, b in zip(cp[1], g.arrays(tag, "f")):
            assert np.linalg.norm(a.numpy() - b) / np.linalg.norm(b) <= 1e-7


def test_sharded_requires_opt_in():
    """sharding never switches on implicitly; sharded=True without a process group is an error."""
    import tensorly_b200 as tb
    from tensorly_b200.cp_als import _Comm
    assert _Comm(None).active is False
    with pytest.raises(RuntimeError):
        _Comm(None, sharded=True)


def test_dimension_tree_sweep_equals_n_pass_sweep_on_cpu_ops():
    """Host logic of the dimension-tree sweep (when T is formed, which modes it serves, when it is dropped),
    with oracle-backed ops in fp64: identical trajectory to the sweep with a full MTTKRP per mode."""
    import tensorly_b200 as tb
    from oracle import oracle as O
    from oracle_ops import OracleOps
    for shape, fixed in (((9, 8, 7), ()), ((6, 5, 4, 7), ()), ((6, 5, 4, 7), (1,))):
        x = torch.from_numpy(O.random_tensor(shape, 0))
        w, fs = O.random_cp_factors(shape, 3, 1)
        runs = []
        for dimtree in (True, False):
            st = tb.CPALS(x, torch.from_numpy(w.copy()), [torch.from_numpy(f.copy()) for f in fs], ops=OracleOps,
                          fixed_modes=fixed, dimtree=dimtree)
            assert st.dimtree == dimtree
            errs = []
            for _ in range(4):
                st.sweep(True)
                errs.append(float(st.err[0]))
                assert st._contracted is None            # T never outlives a sweep
            runs.append((errs, st.factors))
        assert np.allclose(runs[0][0], runs[1][0], rtol=1e-12, atol=0)
        for a, b in zip(runs[0][1], runs[1][1]):
            assert np.allclose(a.numpy(), b.numpy(), rtol=1e-9, atol=1e-12)
    # fewer than two served modes: the reuse cannot pay off and stays off
    x = torch.from_numpy(O.random_tensor((5, 4, 3), 0))
    w, fs = O.random_cp_factors((5, 4, 3), 2, 1)
    st = tb.CPALS(x, torch.from_numpy(w), [torch.from_numpy(f) for f in fs], ops=OracleOps, fixed_modes=(0,), dimtree=True)
    assert st.dimtree is False

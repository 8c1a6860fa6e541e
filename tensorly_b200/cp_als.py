"""CP-ALS drivers that own the sweep loop (single GPU and mode-sharded multi-GPU).

`parafac` / `non_negative_parafac` mirror the reference signatures
(tensorly/decomposition/_cp.py:230, _nn_cp.py:26) for the options on the hot path and
reproduce its arithmetic step by step (_cp.py:394-440, _nn_cp.py:107-152):

    per mode:  V = (w w^T) o prod_{i != mode} F_i^T F_i (+ l2 I)      -> tlb200_cp_update
               M = MTTKRP(X, (w, F), mode)                             -> tlb200_mttkrp
               F_mode = solve(V^T, M^T)^T                              -> tlb200_cp_update
    per sweep: err = sqrt(|‖X‖² + ‖cp‖² − 2<M_last, F_last>|)/‖X‖      -> tlb200_cp_error

Unlike the unmodified reference loop (≈26 array-library calls and one host sync per
mode) a mode update here is 5 kernel launches with no host synchronisation, and the whole
sweep is captured in a CUDA graph.  The unmodified reference drivers also run on the same
kernels through the tenalg backend (tensorly_b200.use()); options that this driver does
not implement (sparsity, linesearch, orthogonalise, normalize_factors) are delegated
to them.  Missing values (`mask`) are handled here: one fused imputation + error pass per sweep.

Multi-GPU (one process per GPU, torch.distributed): the tensor is sharded along
`shard_mode` (each rank holds a contiguous slab and that mode's factor rows); every
other factor is replicated.  Per sweep the ranks exchange one R x R Gram (sharded mode)
and one I_n x R MTTKRP partial per non-sharded mode with all_reduce — NCCL over NVLink on
GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import math
import os
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch

from . import _ops


class CPResult:
    """Minimal (weights, factors) container, iterable like tensorly.cp_tensor.CPTensor."""

    def __init__(self, weights, factors):
        self.weights = weights
        self.factors = list(factors)
        self.rank = self.factors[0].shape[1]
        self.shape = tuple(f.shape[0] for f in self.factors)

    def __iter__(self):
        yield self.weights
        yield self.factors

    def __getitem__(self, i):
        return (self.weights, self.factors)[i]

    def __len__(self):
        return 2


def _wrap(weights, factors):
    try:
        from .backend import import_tensorly
        import_tensorly()
        from tensorly.cp_tensor import CPTensor
        return CPTensor((weights, factors))
    except Exception:  # tensorly absent or backend mismatch: plain container
        return CPResult(weights, factors)


class CudaOps:
    """The product compute path: every call is a tlb200 kernel launch."""

    mttkrp = staticmethod(_ops.mttkrp)
    mode_dot = staticmethod(_ops.mode_dot)
    mttkrp_from_ttm = staticmethod(_ops.mttkrp_from_ttm)   # dimension-tree reuse (see CPALS.dimtree)
    gram = staticmethod(_ops.gram)
    cp_update = staticmethod(_ops.cp_update)
    fused_gram = True        # cp_update(..., gram_out=) also writes the Gram of the updated factor
    # fused right-hand sides: the MTTKRP leaves its split-K partials unsummed, the solve sums them while its LU runs and
    # also returns <M, F_new>, from which the error needs only R x R more work (3 launches and ~35 us less per sweep)
    mttkrp_partials = staticmethod(_ops.mttkrp_partials)
    mttkrp_from_ttm_partials = staticmethod(_ops.mttkrp_from_ttm_partials)
    cp_update_fused = staticmethod(_ops.cp_update_fused)
    cp_error_iprod = staticmethod(_ops.cp_error_iprod)
    range_hint = staticmethod(_ops.RangeHint)
    nncp_update = staticmethod(_ops.nncp_update)
    hals_update = staticmethod(_ops.hals_update)   # the whole HALS inner iteration of one mode in one kernel
    cp_error = staticmethod(_ops.cp_error)
    cp_impute = staticmethod(_ops.cp_impute)     # masked ALS: imputation + both norms in one tensor pass
    sumsq = staticmethod(_ops.sumsq)
    supports_graphs = True


class _Comm:
    """The exchange side of the sharded driver.  Sharding is an explicit opt-in (`sharded=True` or a process
    group): a plain `parafac(x, r)` inside a multi-rank job treats `x` as a whole tensor, never as a slab."""

    def __init__(self, group=None, sharded=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        if sharded is None:
            sharded = group is not None
        ready = dist.is_available() and dist.is_initialized()
        if sharded and not ready:
            raise RuntimeError("sharded=True needs an initialised torch.distributed process group")
        self.active = bool(sharded) and ready and dist.get_world_size(group) > 1
        self.world = dist.get_world_size(group) if self.active else 1
        self.rank = dist.get_rank(group) if self.active else 0
        # GPU ranks of one node: the small partials travel through peer memory (p2p.py, csrc/comm.cu) instead of NCCL;
        # set up lazily on the first CUDA tensor, all ranks agreeing on whether it worked.  TLB200_P2P=0 keeps NCCL.
        self._p2p = None
        self._p2p_tried = False
        self.kind = "torch.distributed all_reduce"

    @property
    def graph_safe(self) -> bool:
        """True when the sweep contains no library collective (captured graphs then own nothing of NCCL's)."""
        return self._p2p is not None

    def _setup_p2p(self, t) -> None:
        self._p2p_tried = True
        if not t.is_cuda or os.environ.get("TLB200_P2P", "1") == "0" or self.dist.get_backend(self.group) != "nccl":
            return
        ok = torch.ones(1, dtype=torch.int32, device=t.device)
        comm = None
        try:
            from .p2p import P2PComm
            comm = P2PComm(self.dist, self.group, t.device)
        except Exception:
            ok.zero_()
        self.dist.all_reduce(ok, op=self.dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 1:
            self._p2p = comm
            self.kind = "one-shot peer-memory all-reduce (own kernel, NVLink stores + flags, rank-ordered sum)"

    def broadcast(self, t, src_rank=0):
        """Make rank `src_rank`'s copy of `t` everyone's (replicated state must be bit-identical)."""
        if self.active:
            src = self.dist.get_global_rank(self.group, src_rank) if self.group is not None else src_rank
            self.dist.broadcast(t, src=src, group=self.group)
        return t

    def all_reduce(self, t):
        if not self.active:
            return t
        if not self._p2p_tried:
            self._setup_p2p(t)
        if self._p2p is not None and self._p2p.fits(t):
            return self._p2p.all_reduce(t)
        self.dist.all_reduce(t, group=self.group)
        return t

    def all_reduce_partials(self, part, tail, out):
        """Fused split-K sum + exchange; only with the peer-memory path (the caller checks `graph_safe`)."""
        return self._p2p.all_reduce_partials(part, tail, out)

    def fits_p2p(self, t) -> bool:
        return self._p2p is not None and self._p2p.fits(t)

    def close(self):
        if self._p2p is not None:
            self._p2p.close()
            self._p2p = None

    def all_gather_rows(self, t):
        """Concatenate per-rank row blocks (equal or unequal row counts)."""
        if not self.active:
            return t
        counts = [torch.zeros(1, dtype=torch.int64, device=t.device) for _ in range(self.world)]
        self.dist.all_gather(counts, torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device), group=self.group)
        counts = [int(c.item()) for c in counts]
        mx = max(counts)
        pad = torch.zeros((mx, t.shape[1]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        parts = [torch.empty_like(pad) for _ in range(self.world)]
        self.dist.all_gather(parts, pad, group=self.group)
        return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def shard_bounds(extent: int, world: int, rank: int):
    """Contiguous block partition of `extent` rows over `world` ranks (first ranks get the
    remainder)."""
    base, rem = divmod(extent, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class CPALS:
    """State + one-sweep step of CP-ALS (least squares or multiplicative updates)."""

    def __init__(self, tensor_local: torch.Tensor, weights: torch.Tensor, factors: Sequence[torch.Tensor],
                 l2_reg: float = 0.0, update: str = "ls", fixed_modes: Sequence[int] = (), comm: Optional[_Comm] = None,
                 shard_mode: int = 0, ops=CudaOps, eps: Optional[float] = None, dimtree: Optional[bool] = None,
                 mask: Optional[torch.Tensor] = None, nn_modes=None, sparsity_coefficients=None, exact: bool = False):
        # update == "hals" (non_negative_parafac_hals, tensorly/decomposition/_nn_cp.py:307-341): modes in `nn_modes`
        # are updated by HALS non-negative least squares, the others by the unconstrained solve
        self.nn_modes = set(range(tensor_local.dim())) if nn_modes in (None, "all") else set(nn_modes)
        self.sparsity_coefficients = list(sparsity_coefficients) if sparsity_coefficients is not None else [None] * tensor_local.dim()
        self.exact = bool(exact)
        # Missing values (mask: 1 = observed, 0 = missing; tensorly/decomposition/_cp.py:195-207, :442-445): the
        # driver owns a copy of the tensor whose missing entries are re-imputed from the current factors after
        # every sweep (before every mode update for the multiplicative rule, _nn_cp.py:124-127).
        self.mask = None
        if mask is not None:
            if tuple(mask.shape) != tuple(tensor_local.shape):
                raise ValueError(f"mask has shape {tuple(mask.shape)} but the tensor has shape {tuple(tensor_local.shape)}")
            self.mask = torch.as_tensor(mask, device=tensor_local.device).to(tensor_local.dtype).contiguous()
            tensor_local = tensor_local.clone(memory_format=torch.contiguous_format)
            self.stats = torch.zeros(3, dtype=tensor_local.dtype, device=tensor_local.device)
        self.x = tensor_local
        self.ops = ops
        self.comm = comm or _Comm()
        self.ndim = tensor_local.dim()
        self.shard_mode = shard_mode if self.comm.active else None
        self.update = update
        self.l2_reg = float(l2_reg or 0.0)
        self.weights = weights
        # all-ones weights (what every driver here produces: the reference keeps the scale in the factors,
        # _cp.py:102-121) are passed to the kernels as "no weights": multiplying by 1 is exact, and the kernels can
        # then read an unweighted factor matrix in place instead of building a weighted Khatri-Rao table first
        self._w = None if bool(torch.all(weights == 1)) else weights
        # own the factor buffers (row-major) — the caller's tensors are never mutated
        self.factors: List[torch.Tensor] = [f.clone().contiguous() for f in factors]
        self.rank = self.factors[0].shape[1]
        for n, f in enumerate(self.factors):
            if f.shape[0] != tensor_local.shape[n]:
                raise ValueError(f"factor {n} has {f.shape[0]} rows but the local tensor has extent "
                                 f"{tensor_local.shape[n]} in mode {n}")
        self.modes = [m for m in range(self.ndim) if m not in set(fixed_modes)]
        self.eps = float(eps if eps is not None else torch.finfo(tensor_local.dtype).eps)
        dt, dev = tensor_local.dtype, tensor_local.device
        self.grams = [torch.empty((self.rank, self.rank), dtype=dt, device=dev) for _ in range(self.ndim)]
        self.err = torch.zeros(3, dtype=dt, device=dev)
        self.norm_x2 = torch.empty(1, dtype=dt, device=dev)
        self.mttkrp_last: Optional[torch.Tensor] = None
        self.iprod = torch.zeros(1, dtype=dt, device=dev)
        self._mbuf = {}
        self._iprod_fresh = False
        # The tensor is constant over the decomposition (no mask): its max |x| is found once and registered, which lets
        # the rank-33..64 tensor passes run on the fp16-split engine (half the tensor-core work per byte).
        self._range_hint = None
        if (self.mask is None and hasattr(self.ops, "range_hint") and tensor_local.is_cuda
                and _ops.RangeHint.applies(tensor_local, self.rank)):
            self._range_hint = self.ops.range_hint(tensor_local)
        self._want_error = True       # set per sweep: lets the last mode's solve finish the error in its own tail
        self._err_done = False
        self._fuse = (update == "ls" and hasattr(self.ops, "cp_update_fused") and tensor_local.is_cuda
                      and os.environ.get("TLB200_FUSED_UPDATE", "1") != "0")
        self._graph = None
        self._graph_key = None
        self._eager_runs = 0
        # Dimension-tree reuse: T = X x_{N-1} F_{N-1}^T is formed once per sweep and serves the MTTKRPs of all
        # modes before the last, so the tensor is streamed twice per sweep instead of N times (same updates, the
        # last factor is only changed after T's last use).  Worth it from two served modes on.
        served = [m for m in self.modes if m < self.ndim - 1]
        can = hasattr(self.ops, "mttkrp_from_ttm") and self.ndim >= 3 and len(served) >= 2
        if dimtree is None:
            dimtree = os.environ.get("TLB200_DIMTREE", "1") != "0"
        self.dimtree = bool(dimtree) and can and not (self.mask is not None and update == "mu")
        self._contracted: Optional[torch.Tensor] = None
        # Sharded runs: the Gram of the sharded mode's new rows (R x R partial) and the MTTKRP partial of the NEXT
        # updated mode are both all-reduced before that mode's solve, so they travel in ONE collective: the Gram
        # lives at the tail of a packed buffer whose head receives that MTTKRP (one sync point less per sweep).
        self._pack = None
        self._pack_mode = None
        if self.comm.active and self.shard_mode in self.modes and getattr(self.ops, "packed_allreduce", True):
            k = self.modes.index(self.shard_mode)
            if k + 1 < len(self.modes):
                pm = self.modes[k + 1]
                rows = tensor_local.shape[pm]
                self._pack = torch.empty(rows * self.rank + self.rank * self.rank, dtype=dt, device=dev)
                self._pack_mode = pm
                self._pack_m = self._pack[: rows * self.rank].view(rows, self.rank)
                self.grams[self.shard_mode] = self._pack[rows * self.rank:].view(self.rank, self.rank)
        # ||X||^2 (all-reduced over slabs) and the initial Grams
        self.ops.sumsq(self.x, out=self.norm_x2)
        self.comm.all_reduce(self.norm_x2)
        for n in range(self.ndim):
            self._refresh_gram(n)

    # -- pieces ---------------------------------------------------------------------------
    def _refresh_gram(self, n: int, defer: bool = False) -> None:
        self.ops.gram(self.factors[n], out=self.grams[n])
        if self.shard_mode == n and not defer:
            self.comm.all_reduce(self.grams[n])

    def _impute(self) -> None:
        """tensor <- tensor * mask + rec * (1 - mask); err <- ||tensor - rec|| / ||tensor||; ||tensor||^2 refreshed."""
        self.ops.cp_impute(self.x, self.mask, (self.weights, self.factors), out=self.x, stats=self.stats)
        if self.shard_mode is not None:
            self.comm.all_reduce(self.stats[1:3])          # both sums are over this rank's slab
            self.stats[0] = torch.sqrt(self.stats[2] / self.stats[1])
        self.err[0].copy_(self.stats[0])
        self.norm_x2.copy_(self.stats[1:2])

    def _update_mode_fused(self, mode: int) -> bool:
        """LS update of a mode whose MTTKRP needs no exchange: partials -> one solve launch (sum + LU + substitution +
        Gram [+ <M, F>]).  Returns False when this problem has no fused form (the caller then takes the plain path)."""
        last = mode == self.ndim - 1
        try:
            if self._contracted is not None and not last:
                part = self.ops.mttkrp_from_ttm_partials(self._contracted, (self._w, self.factors), mode)
            else:
                part = self.ops.mttkrp_partials(self.x, (self._w, self.factors), mode)
            want_ip = last and self.shard_mode != mode
            fold = want_ip and self._want_error and self.mask is None
            self.ops.cp_update_fused(self.grams, mode, self.weights, part, self.l2_reg, out=self.factors[mode],
                                     gram_out=self.grams[mode], iprod_out=self.iprod if want_ip else None,
                                     norm_x2=self.norm_x2 if fold else None, err_out=self.err if fold else None)
        except NotImplementedError:
            self._fuse = False
            return False
        if self.shard_mode == mode and self._pack is None:
            self.comm.all_reduce(self.grams[mode])
        if last:
            self._iprod_fresh = want_ip
            self._err_done = fold
            self.mttkrp_last = None
        return True

    def _update_mode_fused_exchange(self, mode: int) -> bool:
        """LS update of a mode whose MTTKRP is a partial sum over the slabs, on the peer-memory path: the exchange kernel
        sums the split-K partials while it pushes them (no reduction launch), the solve takes the exchanged MTTKRP and
        returns <M, F> for the error."""
        packed = self._pack is not None and mode == self._pack_mode
        rows = self.x.shape[mode]
        if packed:
            buf, m = self._pack, self._pack_m
        else:
            if self._mbuf.get(mode) is None:
                self._mbuf[mode] = torch.empty((rows, self.rank), dtype=self.x.dtype, device=self.x.device)
            buf = m = self._mbuf[mode]
        if not self.comm.fits_p2p(buf):
            return False
        last = mode == self.ndim - 1
        try:
            if self._contracted is not None and not last:
                part = self.ops.mttkrp_from_ttm_partials(self._contracted, (self._w, self.factors), mode)
            else:
                part = self.ops.mttkrp_partials(self.x, (self._w, self.factors), mode)
            self.comm.all_reduce_partials(part, self.grams[self.shard_mode] if packed else None, buf.reshape(-1))
            fold = last and self._want_error and self.mask is None
            self.ops.cp_update_fused(self.grams, mode, self.weights, _ops.plain_partials(m), self.l2_reg,
                                     out=self.factors[mode], gram_out=self.grams[mode],
                                     iprod_out=self.iprod if last else None,
                                     norm_x2=self.norm_x2 if fold else None, err_out=self.err if fold else None)
        except NotImplementedError:
            self._fuse = False
            return False
        if last:
            self._iprod_fresh = True
            self._err_done = fold
            self.mttkrp_last = None
        return True

    def _update_mode(self, mode: int) -> None:
        if self.mask is not None and self.update == "mu":
            self._impute()
        # the fused form applies when this rank's MTTKRP rows are final as they are: single GPU, or the sharded mode
        local = self.shard_mode is None or (self.shard_mode == mode and mode != self.ndim - 1)
        if self._fuse and local and self._update_mode_fused(mode):
            return
        if (self._fuse and not local and self.shard_mode is not None and mode != self.shard_mode and self.comm.graph_safe
                and self._update_mode_fused_exchange(mode)):
            return
        if mode == self.ndim - 1:
            self._iprod_fresh = False
            self._err_done = False
        packed = self._pack is not None and mode == self._pack_mode
        out = self._pack_m if packed else None
        if self._contracted is not None and mode < self.ndim - 1:
            m = self.ops.mttkrp_from_ttm(self._contracted, (self._w, self.factors), mode, out=out)
        else:
            m = self.ops.mttkrp(self.x, (self._w, self.factors), mode, out=out)
        if packed:
            self.comm.all_reduce(self._pack)     # this mode's MTTKRP partial + the sharded mode's Gram partial
        elif self.shard_mode is not None and mode != self.shard_mode:
            self.comm.all_reduce(m)              # partial sums over the slabs
        if self.update == "ls" and getattr(self.ops, "fused_gram", False):
            # one launch: solve + Gram of the new rows (partial over this rank's rows when the mode is sharded)
            self.ops.cp_update(self.grams, mode, self.weights, m, self.l2_reg, out=self.factors[mode],
                               gram_out=self.grams[mode])
            if self.shard_mode == mode and self._pack is None:
                self.comm.all_reduce(self.grams[mode])     # (packed: reduced together with the next mode's MTTKRP)
        else:
            if self.update == "ls" or (self.update == "hals" and mode not in self.nn_modes):
                self.ops.cp_update(self.grams, mode, self.weights, m, self.l2_reg, out=self.factors[mode])
            elif self.update == "hals":
                # hals_nnls(M^T, V, F^T, n_iter_max=100, sparsity_coefficient=..., exact=...) (_nn_cp.py:328-335)
                self.ops.hals_update(self.grams, mode, self.weights, m, self.factors[mode],
                                     n_iter_max=50000 if self.exact else 100, tol=1e-16 if self.exact else 1e-8,
                                     sparsity_coefficient=self.sparsity_coefficients[mode])
            else:
                self.ops.nncp_update(self.grams, mode, self.weights, m, self.factors[mode], self.eps)
            self._refresh_gram(mode, defer=self._pack is not None)   # packed: reduced with the next mode's MTTKRP
        self.mttkrp_last = m

    def _error(self) -> None:
        last = self.ndim - 1
        if self._err_done:          # the last mode's solve already wrote self.err
            return
        if self._iprod_fresh:
            self.ops.cp_error_iprod(self.grams, self.weights, self.iprod, self.norm_x2, out=self.err)
            return
        self.ops.cp_error(self.grams, self.weights, self.mttkrp_last, self.factors[last], self.norm_x2, out=self.err)
        if self.shard_mode == last:
            # <M_last, F_last> was summed over local rows only
            self.comm.all_reduce(self.err[1:2])
            nx2 = self.norm_x2[0]
            self.err[0] = torch.sqrt(torch.abs(nx2 + self.err[2] - 2 * self.err[1])) / torch.sqrt(nx2)

    def sweep_eager(self, with_error: bool = True) -> None:
        self._want_error, self._err_done = bool(with_error), False
        if self.dimtree:
            self._contracted = self.ops.mode_dot(self.x, self.factors[-1], self.ndim - 1, transpose=True)
        for mode in self.modes:
            if mode == self.ndim - 1:
                self._contracted = None       # the last factor changes now: T is stale (and its memory is free again)
            self._update_mode(mode)
        self._contracted = None
        if self.mask is not None:
            self._impute()           # also yields the error (no shortcut through the last MTTKRP with a mask)
            return
        if with_error:
            if self.modes[-1] != self.ndim - 1:
                # the fast error needs the last mode's MTTKRP with the current factors
                self._iprod_fresh = self._err_done = False
                self.mttkrp_last = self.ops.mttkrp(self.x, (self._w, self.factors), self.ndim - 1)
                if self.shard_mode is not None and self.shard_mode != self.ndim - 1:
                    self.comm.all_reduce(self.mttkrp_last)
            self._error()

    def sweep(self, with_error: bool = True, use_graph: bool = True) -> None:
        """One ALS sweep.  On a single GPU the first sweep runs eagerly (lazy one-time
        initialisation), the second is captured into a CUDA graph, and every later sweep
        replays that graph: one launch per sweep instead of ~20."""
        # The sharded sweep is captured too when its exchanges run on the peer-memory kernel (no NCCL inside the
        # graph).  With NCCL collectives it stays eager by default: capturing them works, but process-group teardown
        # then hung in our runs (torch 2.11 / NCCL 2.28); TLB200_DIST_GRAPH=1 opts in.
        graphable = (use_graph and getattr(self.ops, "supports_graphs", False) and self.x.is_cuda
                     and self.update != "hals"          # its kernel is a cooperative launch: kept out of graph capture
                     and (not self.comm.active or self.comm.graph_safe or os.environ.get("TLB200_DIST_GRAPH", "0") == "1")
                     and not (self.comm.active and self.shard_mode == self.ndim - 1))
        if not graphable:
            self.sweep_eager(with_error)
            return
        key = bool(with_error)
        if self._graph_key != key:
            self._graph, self._graph_key, self._eager_runs = None, key, 0
        if self._graph is None:
            if self._eager_runs < 1:
                self.sweep_eager(with_error)
                self._eager_runs += 1
                return
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.sweep_eager(with_error)
            self._graph = g
        self._graph.replay()

    def gathered_factors(self) -> List[torch.Tensor]:
        out = []
        for n, f in enumerate(self.factors):
            out.append(self.comm.all_gather_rows(f) if self.shard_mode == n else f)
        return out


# --------------------------------------------------------------------------------------
def _random_init(shape, rank, random_state, dtype, device, non_negative=False):
    """random_cp(shape, rank, normalise_factors=False) of the reference
    (tensorly/random/base.py:103-114): one RandomState.random_sample((I_n, R)) per mode, in
    mode order; weights = ones."""
    if isinstance(random_state, np.random.RandomState):
        rng = random_state
    else:
        rng = np.random.RandomState(random_state) if random_state is not None else np.random.mtrand._rand
    factors = [torch.as_tensor(rng.random_sample((s, rank)), dtype=dtype).to(device) for s in shape]
    weights = torch.ones(rank, dtype=dtype, device=device)
    return weights, factors


def _nndsvda(matrix, U, S, V):
    """make_svd_non_negative(..., nntype='nndsvda') of the reference (tenalg/svd.py:68-135), left factor only:
    the leading triplet as is, every other column replaced by the dominant of its positive / negative parts
    scaled by sqrt(S_j * |x_+-| * |y_+-|), entries below eps filled with the mean of the matrix.  All columns
    at once, no host round trips (the reference loops over the columns and reads the norms back)."""
    k = min(U.shape[1], V.shape[0])
    x, y = U[:, :k], V[:k, :]
    xp, yp = torch.clamp(x, min=0.0), torch.clamp(y, min=0.0)
    xn, yn = torch.abs(torch.clamp(x, max=0.0)), torch.abs(torch.clamp(y, max=0.0))
    xpn, ypn = torch.linalg.norm(xp, dim=0), torch.linalg.norm(yp, dim=1)
    xnn, ynn = torch.linalg.norm(xn, dim=0), torch.linalg.norm(yn, dim=1)
    m_p, m_n = xpn * ypn, xnn * ynn
    pos = m_p > m_n
    u = torch.where(pos, xp / xpn, xn / xnn)
    lbd = torch.sqrt(S[:k] * torch.where(pos, m_p, m_n))
    W = torch.zeros_like(U)
    W[:, :k] = u * lbd
    W[:, 0] = torch.sqrt(S[0]) * torch.abs(U[:, 0])
    eps = torch.finfo(matrix.dtype).eps
    return torch.where(W < eps, matrix.mean().expand_as(W), W)


def _svd_init(tensor, rank, random_state, non_negative=False):
    """init='svd' of the reference (_cp.py:72-100 through svd_interface, tenalg/svd.py:366-447): leading left
    singular vectors of each unfolding (torch.linalg.svd on our unfold), sign-fixed by svd_flip, made
    non-negative by NNDSVDA when asked, mode-0 vectors scaled by the singular values, random padding when
    I_n < rank."""
    rng = np.random.RandomState(random_state) if not isinstance(random_state, np.random.RandomState) else random_state
    factors = []
    for mode in range(tensor.dim()):
        unf = _ops.unfold(tensor, mode)
        U, S, V = torch.linalg.svd(unf, full_matrices=False)
        # svd_flip (tenalg/svd.py:13-40): largest-|.| entry of each column positive
        idx = torch.argmax(torch.abs(U), dim=0)
        signs = torch.sign(U[idx, torch.arange(U.shape[1], device=U.device)])
        U = U * signs
        U, S = U[:, :rank], S[:rank]              # truncated_svd(n_eigenvecs=rank), tenalg/svd.py:232-235
        if non_negative:
            U = _nndsvda(unf, U, S, (V[:U.shape[1]] * signs[:U.shape[1], None])[:rank])
        U = U.clone()
        if mode == 0:
            k = min(rank, S.shape[0])
            U[:, :k] = U[:, :k] * S[:k]
        if tensor.shape[mode] < rank:
            extra = torch.as_tensor(rng.random_sample((U.shape[0], rank - tensor.shape[mode])), dtype=tensor.dtype)
            U = torch.cat([U, extra.to(tensor.device)], dim=1)
        factors.append(U[:, :rank].contiguous())
    weights = torch.ones(rank, dtype=tensor.dtype, device=tensor.device)
    return weights, factors


def _init_from(init, tensor, rank):
    """init=(weights, factors) / CPTensor (_cp.py:102-121): weights other than ones are
    spread evenly over the factors; the caller's tensors are not mutated."""
    weights, factors = init
    factors = [torch.as_tensor(f, dtype=tensor.dtype, device=tensor.device) for f in factors]
    if weights is not None:
        w = torch.as_tensor(weights, dtype=tensor.dtype, device=tensor.device)
        if not bool(torch.all(w == 1)):
            avg = torch.prod(w) ** (1.0 / w.shape[0])
            factors = [f * avg for f in factors]
    return torch.ones(rank, dtype=tensor.dtype, device=tensor.device), factors


def _reference_init(tensor, rank, svd, non_negative, random_state, mask, svd_mask_repeats):
    """initialize_cp of the reference (tensorly/decomposition/_cp.py:26-152) on the pytorch backend."""
    from .backend import import_tensorly, use
    tl = import_tensorly()
    if tl.get_backend() != "pytorch":
        tl.set_backend("pytorch")
    use()
    from tensorly.decomposition._cp import initialize_cp
    kt = initialize_cp(tensor, rank, init="svd", svd=svd, non_negative=non_negative, random_state=random_state,
                       mask=torch.as_tensor(mask, device=tensor.device).to(tensor.dtype), svd_mask_repeats=svd_mask_repeats)
    return torch.ones(rank, dtype=tensor.dtype, device=tensor.device), [f.contiguous() for f in kt.factors]


def _delegate(name, tensor, rank, kwargs):
    from .backend import import_tensorly, use
    tl = import_tensorly()
    if tl.get_backend() != "pytorch":
        tl.set_backend("pytorch")
    use()
    from tensorly import decomposition
    return getattr(decomposition, name)(tensor, rank, **kwargs)


def _run(tensor, rank, n_iter_max, init, svd, tol, random_state, verbose, return_errors, l2_reg, cvg_criterion,
         fixed_modes, callback, update, group, shard_mode, ops, use_graph, sharded=None, mask=None,
         svd_mask_repeats=5, hals=None):
    if not isinstance(tensor, torch.Tensor):
        raise TypeError("tensor must be a torch.Tensor")
    comm = _Comm(group, sharded)
    ndim = tensor.dim()
    rank = int(rank)
    # global shape: the local slab's extent along shard_mode is summed over ranks
    shape = list(tensor.shape)
    lo = 0
    if comm.active:
        ext = torch.tensor([tensor.shape[shard_mode]], dtype=torch.int64, device=tensor.device)
        exts = [torch.zeros_like(ext) for _ in range(comm.world)]
        comm.dist.all_gather(exts, ext, group=comm.group)
        exts = [int(e.item()) for e in exts]
        shape[shard_mode] = sum(exts)
        lo = sum(exts[: comm.rank])
    non_negative = update in ("mu", "hals")
    if isinstance(init, str):
        if init == "random":
            weights, factors = _random_init(shape, rank, random_state, tensor.dtype, tensor.device)
            if comm.active and not isinstance(random_state, int):
                for f in factors:                 # unseeded / stateful generators differ per rank
                    comm.broadcast(f)
        elif init == "svd":
            if comm.active:
                raise NotImplementedError("init='svd' is not available for a sharded tensor; pass init='random' or factors")
            if mask is not None:
                # SVD of an incomplete tensor: the reference imputes inside svd_interface (tenalg/svd.py:431-439);
                # initialisation is not on the hot path, so it is the reference's own routine that runs
                weights, factors = _reference_init(tensor, rank, svd, non_negative, random_state, mask, svd_mask_repeats)
            else:
                weights, factors = _svd_init(tensor, rank, random_state, non_negative=non_negative)
            if non_negative:
                factors = [torch.abs(f) for f in factors]
        else:
            raise ValueError(f'Initialization method "{init}" not recognized')
    else:
        weights, factors = _init_from(init, tensor, rank)
    if comm.active:
        hi = lo + tensor.shape[shard_mode]
        if factors[shard_mode].shape[0] == shape[shard_mode]:
            factors[shard_mode] = factors[shard_mode][lo:hi]
    fixed_modes = list(fixed_modes or [])
    if ndim - 1 in fixed_modes:
        import warnings
        warnings.warn("You asked for fixing the last mode, which is not supported.\n The last mode will not be fixed. "
                      "Consider using tl.moveaxis()")
        fixed_modes.remove(ndim - 1)
    state = CPALS(tensor, weights, factors, l2_reg=l2_reg, update=update, fixed_modes=fixed_modes, comm=comm,
                  shard_mode=shard_mode, ops=ops, mask=mask, **(hals or {}))
    want_err = bool(tol) or return_errors
    err_hist = torch.zeros(max(n_iter_max, 1), dtype=tensor.dtype, device=tensor.device)
    rec_errors: List[float] = []
    done = 0
    for it in range(n_iter_max):
        if verbose > 1:
            print("Starting iteration", it + 1)
        state.sweep(with_error=want_err, use_graph=use_graph)
        done = it + 1
        if want_err:
            err_hist[it] = state.err[0]
        if callback is not None:
            cp_now = _wrap(state.weights, state.gathered_factors())
            if callback(cp_now, float(err_hist[it]) if want_err else None) is True:
                break
        if tol and it >= 1:
            e = err_hist[it - 1: it + 1].tolist()     # one host read per sweep, like the reference
            dec = e[0] - e[1]
            if verbose:
                print(f"iteration {it}, reconstruction error: {e[1]}, decrease = {dec}")
            if cvg_criterion == "abs_rec_error":
                stop = abs(dec) < tol
            elif cvg_criterion == "rec_error":
                stop = dec < tol
            else:
                raise TypeError("Unknown convergence criterion")
            if stop:
                if verbose:
                    print(f"PARAFAC converged after {it} iterations")
                break
    if want_err:
        rec_errors = err_hist[:done].tolist()
    cp = _wrap(state.weights, state.gathered_factors())
    if return_errors:
        return cp, rec_errors
    return cp


def parafac(tensor, rank, n_iter_max=100, init="svd", svd="truncated_svd", normalize_factors=False,
            orthogonalise=False, tol=1e-8, random_state=None, verbose=0, return_errors=False, sparsity=None,
            l2_reg=0, mask=None, cvg_criterion="abs_rec_error", fixed_modes=None, svd_mask_repeats=5,
            linesearch=False, callback=None, *, sharded=None, group=None, shard_mode=0, ops=CudaOps, use_graph=True):
    """CANDECOMP/PARAFAC by ALS — same signature and semantics as
    tensorly.decomposition.parafac (tensorly/decomposition/_cp.py:230).

    Keyword-only extras: `sharded=True` (or a process `group`) runs the sharded multi-GPU algorithm — pass the
    LOCAL slab of the tensor along `shard_mode`; without either the tensor is decomposed on this GPU alone, also
    inside a multi-rank job.  `use_graph=False` disables CUDA graphs.
    """
    if normalize_factors or orthogonalise or sparsity or linesearch or svd != "truncated_svd":
        if _Comm(group, sharded).active:
            raise NotImplementedError("normalize_factors/orthogonalise/sparsity/linesearch are not available sharded")
        return _delegate("parafac", tensor, rank, dict(
            n_iter_max=n_iter_max, init=init, svd=svd, normalize_factors=normalize_factors, orthogonalise=orthogonalise,
            tol=tol, random_state=random_state, verbose=verbose, return_errors=return_errors, sparsity=sparsity,
            l2_reg=l2_reg, mask=mask, cvg_criterion=cvg_criterion, fixed_modes=fixed_modes,
            svd_mask_repeats=svd_mask_repeats, linesearch=linesearch, callback=callback))
    return _run(tensor, rank, n_iter_max, init, svd, tol, random_state, verbose, return_errors, l2_reg, cvg_criterion,
                fixed_modes, callback, "ls", group, shard_mode, ops, use_graph, sharded, mask, svd_mask_repeats)


def non_negative_parafac(tensor, rank, n_iter_max=100, init="svd", svd="truncated_svd", tol=10e-7, random_state=None,
                         verbose=0, normalize_factors=False, return_errors=False, mask=None,
                         cvg_criterion="abs_rec_error", fixed_modes=None, *, sharded=None, group=None, shard_mode=0,
                         ops=CudaOps, use_graph=True):
    """Non-negative CP by multiplicative updates — same signature and semantics as
    tensorly.decomposition.non_negative_parafac (tensorly/decomposition/_nn_cp.py:26)."""
    if normalize_factors or svd != "truncated_svd":
        if _Comm(group, sharded).active:
            raise NotImplementedError("normalize_factors is not available sharded")
        return _delegate("non_negative_parafac", tensor, rank, dict(
            n_iter_max=n_iter_max, init=init, svd=svd, tol=tol, random_state=random_state, verbose=verbose,
            normalize_factors=normalize_factors, return_errors=return_errors, mask=mask, cvg_criterion=cvg_criterion,
            fixed_modes=fixed_modes))
    return _run(tensor, rank, n_iter_max, init, svd, tol, random_state, verbose, return_errors, 0.0, cvg_criterion,
                fixed_modes, None, "mu", group, shard_mode, ops, use_graph, sharded, mask)


def non_negative_parafac_hals(tensor, rank, n_iter_max=100, init="svd", svd="truncated_svd", tol=10e-8, random_state=None,
                              sparsity_coefficients=None, fixed_modes=None, nn_modes="all", exact=False,
                              normalize_factors=False, verbose=False, return_errors=False, cvg_criterion="abs_rec_error", *,
                              sharded=None, group=None, shard_mode=0, ops=CudaOps):
    """Non-negative CP by HALS — same signature and semantics as tensorly.decomposition.non_negative_parafac_hals
    (tensorly/decomposition/_nn_cp.py:186-379).  Per mode: Gram-Hadamard, MTTKRP, then the whole hals_nnls inner
    iteration (up to 100 Gauss-Seidel passes over the rank, tensorly/solvers/nnls.py:139-173) as ONE kernel
    (tlb200_hals_update) instead of ~6 array-library calls per rank and pass."""
    if normalize_factors or svd != "truncated_svd":
        if _Comm(group, sharded).active:
            raise NotImplementedError("normalize_factors is not available sharded")
        return _delegate("non_negative_parafac_hals", tensor, rank, dict(
            n_iter_max=n_iter_max, init=init, svd=svd, tol=tol, random_state=random_state,
            sparsity_coefficients=sparsity_coefficients, fixed_modes=fixed_modes, nn_modes=nn_modes, exact=exact,
            normalize_factors=normalize_factors, verbose=verbose, return_errors=return_errors, cvg_criterion=cvg_criterion))
    n_modes = tensor.dim()
    if sparsity_coefficients is None or isinstance(sparsity_coefficients, float):
        sparsity_coefficients = [sparsity_coefficients] * n_modes
    sparsity_coefficients = list(sparsity_coefficients)
    fixed = list(fixed_modes or [])
    for m in fixed:
        sparsity_coefficients[m] = None
    nn = set(range(n_modes)) if nn_modes == "all" else set(nn_modes or ())
    for m in range(n_modes):
        if sparsity_coefficients[m] is not None and m not in nn:
            import warnings
            warnings.warn("Sparsity coefficient is ignored in unconstrained modes.")
    # the reference evaluates the error only when tol is truthy (_nn_cp.py:342) and never fixes... any mode may be fixed
    out = _run(tensor, rank, n_iter_max, init, svd, tol, random_state, verbose, True, 0.0, cvg_criterion, fixed, None,
               "hals", group, shard_mode, ops, False, sharded, None, 5,
               dict(nn_modes=nn, sparsity_coefficients=sparsity_coefficients, exact=exact))
    cp, errs = out
    if not tol:
        errs = []
    return (cp, errs) if return_errors else cp

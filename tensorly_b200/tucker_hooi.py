"""Tucker decomposition by HOOI with every in-loop step on the hand-written kernels.

`tucker` / `partial_tucker` mirror the reference signatures (tensorly/decomposition/_tucker.py:224 / :105) and the
loop of :187-216:

    per mode k :  Y = multi_mode_dot(X, factors, skip=k, transpose=True)        -> tlb200_multi_mode_dot (tcgen05)
                  factors[k] = rank[k] leading left singular vectors of unfold(Y, k)
    per sweep  :  core = multi_mode_dot(X, factors, transpose=True);  err = sqrt(|‖X‖² − ‖core‖²|) / ‖X‖

The reference takes the singular vectors from a full LAPACK / cuSOLVER SVD of the (I_k x prod of the other ranks)
unfolding — 98 % of a C3 sweep once the TTM chains run on the tensor cores.  HOOI, however, only needs an
orthonormal basis of the dominant left subspace, which is also the dominant eigenspace of the small Gram matrix
G = unfold(Y, k) unfold(Y, k)^T (I_k x I_k).  Here:

    G   = mode_dot(unfold(Y, k), unfold(Y, k), 1)             one pass of the TTM engine (3xTF32), fp32
    U  <- orth(G U)   `svd_iters` times, warm-started from the previous sweep's factors[k]: tlb200_subspace_iterate,
                      two launches per step (fp64 GEMM block + Gram partial + Cholesky-QR factor; apply)

The iteration runs in fp64 on the tiny matrices because a tensor with a dominant mean component makes G U
ill-conditioned beyond fp32 after one step.  No SVD, no eigendecomposition, no host synchronisation inside the loop.

What is (deliberately) different from the reference: the factors span the same subspaces but are not the singular
vectors themselves (any orthonormal basis gives the same reconstruction — Tucker factors are only defined up to a
rotation that the core absorbs), and the subspace is converged by a fixed number of power steps instead of to LAPACK
precision: the reconstruction-error trajectory agrees with the reference's to < 1e-4 relative with the default
`svd_iters=6` (tests/test_gpu_parity.py), tighter with more.  `init="svd"` uses one library `eigh` per mode on the
Gram matrix of the raw unfolding — initialisation, outside the loop.  Options outside this path (mask,
fixed_factors, a non-default `svd`, rank > 64) are delegated to the unmodified reference driver on the b200 backend.
"""
from __future__ import annotations

import math
import warnings
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _ops

MAX_RANK = 64          # widest block tlb200_orthonormalize takes


def _delegate(name, *args, **kwargs):
    from .backend import import_tensorly, use
    from .svd import use_gram_svd
    tl = import_tensorly()
    if tl.get_backend() != "pytorch":
        tl.set_backend("pytorch")
    use()
    use_gram_svd()
    from tensorly import decomposition
    return getattr(decomposition, name)(*args, **kwargs)


class CudaOps:
    """The product compute path: every call is a tlb200 kernel launch."""

    multi_mode_dot = staticmethod(_ops.multi_mode_dot)
    mode_dot = staticmethod(_ops.mode_dot)
    unfold = staticmethod(_ops.unfold)
    orthonormalize = staticmethod(_ops.orthonormalize)
    symeig = staticmethod(_ops.symeig)
    subspace_iterate = staticmethod(_ops.subspace_iterate)
    sumsq = staticmethod(_ops.sumsq)
    supports_graphs = True
    range_hint = staticmethod(_ops.RangeHint)

    @staticmethod
    def eigh_top(g, p):
        """Leading p eigenvectors of a symmetric fp64 matrix by a library eigendecomposition — used ONLY while the
        projections still run on factors that are not orthonormal (the first sweep of a random / user-given init,
        and init="svd" itself), i.e. as part of the initialisation; the steady-state loop never calls it."""
        _, vec = torch.linalg.eigh(g)
        return torch.flip(vec[:, -int(p):], dims=(1,)).contiguous()


def _gram_of_unfolding(ops, y: torch.Tensor, mode: int, double: bool = False) -> torch.Tensor:
    """unfold(y, mode) unfold(y, mode)^T as ONE call of the TTM kernel: contracting the unfolding's long side with
    itself, out[i, i'] = sum_c U[i', c] U[i, c].  `double`: form it in fp64 (SIMT path) — squaring halves the dynamic
    range, and an fp32 Gram matrix loses every direction whose singular value is below 2.4e-4 of the largest."""
    unf = ops.unfold(y, mode, contiguous=True) if mode != 0 else y.reshape(y.shape[0], -1)
    if double and unf.dtype != torch.float64:
        unf = unf.to(torch.float64)
    return ops.mode_dot(unf, unf, 1)


class HOOI:
    """State + one sweep of HOOI over `modes` (all other modes are left untouched: partial Tucker)."""

    def __init__(self, tensor: torch.Tensor, rank: Sequence[int], modes: Sequence[int], factors: Sequence[torch.Tensor],
                 svd_iters: int = 6, ops=CudaOps):
        self.ops = ops
        self.x = tensor if tensor.is_contiguous() else tensor.contiguous()
        self.modes = list(modes)
        self.rank = [int(r) for r in rank]
        self.svd_iters = int(svd_iters)
        # Per mode a block of p >= r orthonormal vectors in fp64 (the power iteration runs on it):
        #   I_k <= 64          p = I_k: the block spans the whole space, the Rayleigh-Ritz step alone is exact
        #   otherwise          p = min(64, 2 r): oversampling — the wanted r vectors converge like
        #                      (lambda_{p+1} / lambda_r)^steps instead of (lambda_{r+1} / lambda_r)^steps
        self.block: List[Optional[torch.Tensor]] = []
        self.factors: List[torch.Tensor] = []
        gen = torch.Generator(device=self.x.device).manual_seed(0x5eed) if self.x.is_cuda else torch.Generator().manual_seed(0x5eed)
        for f, r, m in zip(factors, self.rank, self.modes):
            extent = self.x.shape[m]
            if tuple(f.shape) != (extent, r):
                raise ValueError(f"factor of mode {m} must be ({extent}, {r}), got {tuple(f.shape)}")
            f64 = f.to(torch.float64)
            if extent <= MAX_RANK:
                self.block.append(None)                               # exact path: no block to carry
            else:
                p = min(MAX_RANK, 2 * r)
                pad = torch.rand((extent, p - r), generator=gen, dtype=torch.float64, device=self.x.device) - 0.5
                self.block.append(ops.orthonormalize(torch.cat([f64, pad], dim=1).contiguous()))
            # the projections of the first sweep use the initial factors AS GIVEN (a random init is not
            # orthonormal), exactly like the reference loop (_tucker.py:194-196)
            self.factors.append(f.to(self.x.dtype).contiguous())
        self.norm_x2 = ops.sumsq(self.x)
        # the first product of every chain reads the (constant) input tensor: with its max |x| registered, projections
        # onto more than 32 columns run on the fp16-split engine (tensorly_b200._ops.RangeHint)
        self._range_hint = None
        if self.x.is_cuda and hasattr(ops, "range_hint") and _ops.RangeHint.applies(self.x, max(int(r) for r in rank)):
            self._range_hint = ops.range_hint(self.x)
        self.core: Optional[torch.Tensor] = None
        self.err = torch.zeros(1, dtype=self.x.dtype, device=self.x.device)
        self._graph = None
        self._eager_runs = 0
        self._stable = False
        self._orthonormal = False        # becomes True once every factor has been replaced by an orthonormal basis

    def _set_factor(self, index: int, value: torch.Tensor) -> None:
        if self._stable:
            self.factors[index].copy_(value)
        else:
            self.factors[index] = value.to(self.x.dtype).contiguous()

    def _update(self, index: int) -> None:
        mode, r = self.modes[index], self.rank[index]
        ops = self.ops
        y = ops.multi_mode_dot(self.x, self.factors, modes=self.modes, skip=index, transpose=True)
        # While the projections still use factors that are not orthonormal (the first sweep of a random init), Y is
        # nearly rank one — its interesting directions sit 1e-3..1e-4 below the leading singular value, beyond what an
        # fp32 Gram matrix resolves — so that sweep forms G in fp64; afterwards the tcgen05 engine (3xTF32) does.
        g = _gram_of_unfolding(ops, y, mode, double=not self._orthonormal).to(torch.float64)
        u = self.block[index]
        if u is None:
            # small mode: the eigenvectors of G itself, exactly the reference's singular vectors (up to sign)
            _, vec = ops.symeig(g)
            self._set_factor(index, vec[:, :r])
            return
        if not self._orthonormal:
            # First sweep on factors that are not orthonormal (a random init is all-positive: Y is almost rank one,
            # lambda_1 / lambda_2 of G ~ 1e8 and beyond).  Power steps with a Cholesky-QR cannot hold on to the
            # subdominant directions there (measured at C3: the more steps, the worse the first sweep), so this
            # one sweep takes the exact eigenvectors of the fp64 Gram matrix and hands the iteration a proper start.
            u = ops.eigh_top(g, u.shape[1])
            self.block[index].copy_(u)
            self._set_factor(index, u[:, :r])
            return
        # without room to oversample (r already at the 64-column limit) the wanted vectors converge at the slower
        # (lambda_{r+1} / lambda_r) rate: twice the steps
        steps = self.svd_iters if u.shape[1] > r else 2 * self.svd_iters
        u = ops.subspace_iterate(g, u, steps)                    # U <- orth(G U), `steps` times, in place
        u = ops.orthonormalize(u, out=u, passes=2)               # the steps use a shifted Cholesky-QR: clean up
        if u.shape[1] > r:
            # Rayleigh-Ritz: rotate the block so that its first r columns are the leading Ritz vectors
            z = ops.mode_dot(g, u, 1, transpose=True)
            h = ops.mode_dot(z, u, 0, transpose=True)           # U^T G U  (p x p)
            _, w = ops.symeig(h)
            u = ops.mode_dot(u, w, 1, transpose=True)           # U W
        if u is not self.block[index]:
            self.block[index].copy_(u)
        self._set_factor(index, u[:, :r])

    def sweep_eager(self) -> None:
        for index in range(len(self.modes)):
            self._update(index)
        self._orthonormal = True
        core = self.ops.multi_mode_dot(self.x, self.factors, modes=self.modes, transpose=True)
        nc2 = self.ops.sumsq(core)
        err = torch.sqrt(torch.abs(self.norm_x2 - nc2)) / torch.sqrt(self.norm_x2)
        if self.core is None or self.core.shape != core.shape:
            self.core = core
        else:
            self.core.copy_(core)          # stable buffers: a captured sweep keeps writing the same tensors
        self.err.copy_(err)

    def sweep(self, use_graph: bool = True) -> None:
        """One HOOI sweep.  On the GPU the first sweep runs eagerly (it replaces the initial factors, which may
        be the caller's), the second is captured into a CUDA graph and later sweeps replay it: ~170 small launches
        per sweep with no host work in between."""
        graphable = use_graph and self.x.is_cuda and getattr(self.ops, "supports_graphs", False)
        if not graphable:
            self.sweep_eager()
            return
        if self._graph is None:
            if self._eager_runs < 1:
                self.sweep_eager()
                self._stabilise()
                self._eager_runs += 1
                return
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.sweep_eager()
            self._graph = g
        self._graph.replay()

    def _stabilise(self) -> None:
        """Give factors and blocks buffers of their own that every later sweep updates in place."""
        self.factors = [f.clone() for f in self.factors]
        self._stable = True


def _svd_init(ops, x: torch.Tensor, rank, modes):
    """init='svd' (initialize_tucker, _tucker.py:63-77): leading left singular vectors of every raw unfolding, from
    the eigendecomposition of its Gram matrix (one TTM-kernel pass over X per mode + a library eigh of an
    I_k x I_k matrix — initialisation only)."""
    factors = []
    for r, m in zip(rank, modes):
        # fp64 Gram matrix when its fp64 copy of the unfolding is affordable (<= 2 GiB), else the tensor-core one
        g = _gram_of_unfolding(ops, x, m, double=x.numel() <= (1 << 28))
        factors.append(ops.eigh_top(g.to(torch.float64), r))
    return factors


def _random_init(x: torch.Tensor, rank, modes, random_state):
    """init='random' (_tucker.py:81-93): the core is drawn first, then one random_sample((I_k, r_k)) per mode."""
    rng = random_state if isinstance(random_state, np.random.RandomState) else (
        np.random.RandomState(random_state) if random_state is not None else np.random.mtrand._rand)
    rng.random_sample([int(r) for r in rank])
    return [torch.as_tensor(rng.random_sample((x.shape[m], int(r)))).to(x.device) for r, m in zip(rank, modes)]


def partial_tucker(tensor, rank, modes=None, n_iter_max=100, init="svd", tol=10e-5, svd="truncated_svd", random_state=None,
                   verbose=False, mask=None, svd_mask_repeats=5, *, svd_iters=6, ops=CudaOps):
    """Partial Tucker decomposition via HOOI — same signature and return value ((core, factors), rec_errors) as
    tensorly.decomposition.partial_tucker (tensorly/decomposition/_tucker.py:105-221).  `svd_iters` (keyword-only)
    is the number of warm-started power steps that stand in for the reference's SVD per mode and sweep."""
    if not isinstance(tensor, torch.Tensor):
        raise TypeError("tensor must be a torch.Tensor")
    if modes is None:
        modes = list(range(tensor.dim()))
    modes = [int(m) for m in modes]
    if rank is None:
        warnings.warn("No value given for 'rank'. The decomposition will preserve the original size.", Warning)
        rank = [tensor.shape[m] for m in modes]
    elif isinstance(rank, int):
        warnings.warn(f"Given only one int for 'rank' instead of a list of {len(modes)} modes. Using this rank for all modes.",
                      Warning)
        rank = tuple(rank for _ in modes)
    else:
        rank = tuple(int(r) for r in rank)
    too_wide = any(r > MAX_RANK or r > tensor.shape[m] for r, m in zip(rank, modes))
    if mask is not None or svd != "truncated_svd" or too_wide or not (isinstance(init, str) or len(init) == 2):
        return _delegate("partial_tucker", tensor, rank, modes=modes, n_iter_max=n_iter_max, init=init, tol=tol, svd=svd,
                         random_state=random_state, verbose=verbose, mask=mask, svd_mask_repeats=svd_mask_repeats)
    x = tensor if tensor.is_contiguous() else tensor.contiguous()
    if isinstance(init, str):
        if init == "svd":
            factors = _svd_init(ops, x, rank, modes)
        elif init == "random":
            factors = _random_init(x, rank, modes, random_state)
        else:
            raise ValueError(f'Initialization method "{init}" not recognized')
    else:
        _, factors = init
        factors = [torch.as_tensor(f, device=x.device) for f in factors]
    state = HOOI(x, rank, modes, factors, svd_iters=svd_iters, ops=ops)
    if isinstance(init, str) and init == "svd":
        state._orthonormal = True            # eigenvectors of the raw unfoldings' Gram matrices: orthonormal already
    errs = torch.zeros(max(n_iter_max, 1), dtype=x.dtype, device=x.device)
    rec_errors: List[float] = []
    done = 0
    for it in range(n_iter_max):
        state.sweep()
        errs[it] = state.err[0]
        done = it + 1
        if it > 1 and tol:
            e = errs[it - 1: it + 1].tolist()           # one host read per sweep, like the reference
            if verbose:
                print(f"reconstruction error={e[1]}, variation={e[0] - e[1]}.")
            if abs(e[0] - e[1]) < tol:
                if verbose:
                    print(f"converged in {it} iterations.")
                break
    rec_errors = errs[:done].tolist()
    if state.core is None:                              # n_iter_max == 0
        state.core = ops.multi_mode_dot(x, state.factors, modes=modes, transpose=True)
    return (state.core, list(state.factors)), rec_errors


def tucker(tensor, rank, fixed_factors=None, n_iter_max=100, init="svd", return_errors=False, svd="truncated_svd", tol=10e-5,
           random_state=None, mask=None, verbose=False, *, svd_iters=6, ops=CudaOps):
    """Tucker decomposition via HOOI — same signature as tensorly.decomposition.tucker
    (tensorly/decomposition/_tucker.py:224-345); returns a TuckerTensor (or the plain (core, factors) pair when
    TensorLy is not importable), plus the error list with return_errors=True."""
    if fixed_factors:
        return _delegate("tucker", tensor, rank, fixed_factors=fixed_factors, n_iter_max=n_iter_max, init=init,
                         return_errors=return_errors, svd=svd, tol=tol, random_state=random_state, mask=mask, verbose=verbose)
    if not isinstance(tensor, torch.Tensor):
        raise TypeError("tensor must be a torch.Tensor")
    modes = list(range(tensor.dim()))
    if isinstance(rank, int):
        rank = [min(rank, s) for s in tensor.shape]      # validate_tucker_rank for an int: the same rank everywhere
    (core, factors), errs = partial_tucker(tensor, rank, modes, n_iter_max=n_iter_max, init=init, tol=tol, svd=svd,
                                           random_state=random_state, verbose=verbose, mask=mask, svd_iters=svd_iters,
                                           ops=ops)
    try:
        from .backend import import_tensorly
        import_tensorly()
        from tensorly.tucker_tensor import TuckerTensor
        out = TuckerTensor((core, factors))
    except Exception:
        out = (core, factors)
    if return_errors:
        return out, errs
    return out


def tucker_to_tensor(tucker_tensor, skip_factor=None, transpose_factors=False, modes=None):
    """Full tensor of a Tucker decomposition (tensorly/tucker_tensor.py:50-75): one fused chain of TTMs."""
    core, factors = tucker_tensor
    return _ops.multi_mode_dot(core, factors, skip=skip_factor, transpose=transpose_factors, modes=modes)

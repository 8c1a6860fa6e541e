"""Host-side mirror of the reference's tenalg interface for the hot path.

Same names, argument meaning and error behaviour as
  tensorly.base.unfold / fold                                  (tensorly/base.py:39-79)
  tensorly.tenalg.core_tenalg.khatri_rao                       (_khatri_rao.py:9-109)
  tensorly.tenalg.core_tenalg.unfolding_dot_khatri_rao         (mttkrp.py:9-50)
  tensorly.tenalg.core_tenalg.mode_dot / multi_mode_dot        (n_mode_product.py:5-135)
but every arithmetic step runs in the hand-written sm_100a kernels behind the C ABI of
include/tlb200.h.  torch is used only for device memory and the current stream.
Inputs are borrowed and never mutated; results are new tensors.
"""
from __future__ import annotations

import os
import warnings
import weakref
from typing import Iterable, Sequence

import torch

from . import _lib

_DTYPES = {torch.float32: _lib.F32, torch.float64: _lib.F64}

# Kernel family used by MTTKRP / TTM: "auto" (tcgen05 when eligible, SIMT otherwise),
# "simt" or "tcgen05".  Both are CUDA; there is no host fallback.
_path = "auto"


def set_kernel_path(path: str) -> None:
    if path not in _lib.PATHS:
        raise ValueError(f"unknown kernel path {path!r}; expected one of {sorted(_lib.PATHS)}")
    global _path
    _path = path


def get_kernel_path() -> str:
    return _path


def last_kernel_path() -> str:
    """Kernel family the last call on this thread dispatched to ("simt", "tcgen05", ...)."""
    return _lib.load().tlb200_last_path().decode()


def launch_count() -> int:
    """Kernel launches issued by libtlb200 so far (host-side count)."""
    return int(_lib.load().tlb200_launch_count())


# --------------------------------------------------------------------------- helpers
def _check_tensor(t, name: str, ref: torch.Tensor | None = None) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor on a CUDA device (got {type(t).__name__}); "
                        "use tl.set_backend('pytorch') and CUDA tensors with the b200 tenalg backend")
    if not t.is_cuda:
        raise TypeError(f"{name} must live on a CUDA device (got {t.device}); the b200 backend has no CPU path")
    if t.dtype not in _DTYPES:
        raise TypeError(f"{name} has dtype {t.dtype}; the b200 backend supports float32 and float64")
    if t.requires_grad:
        raise TypeError(f"{name} requires grad; the b200 kernels are not differentiable")
    if ref is not None:
        if t.device != ref.device:
            raise TypeError(f"{name} is on {t.device} but the tensor is on {ref.device}")
        if t.dtype != ref.dtype:
            raise TypeError(f"{name} has dtype {t.dtype} but the tensor has dtype {ref.dtype}")
    return t


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


class _Device:
    """Make the tensor's device current for the launch (no-op when it already is)."""

    def __init__(self, t: torch.Tensor):
        self.idx = t.device.index if t.device.index is not None else torch.cuda.current_device()
        self.prev = None

    def __enter__(self):
        cur = torch.cuda.current_device()
        if cur != self.idx:
            self.prev = cur
            torch.cuda.set_device(self.idx)

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)


# Scratch buffers are cached per (device, stream) and only ever grow: calls on one stream are serialised, so
# reusing the buffer is safe, and it keeps 100 MB-sized cudaMalloc/cudaFree pairs out of the hot loop.
# (Under CUDA-graph capture the capture stream is a different key, so the buffer comes from the graph's pool.)
_WS_CACHE: dict = {}


def _workspace(nbytes: int, like: torch.Tensor) -> torch.Tensor:
    nbytes = max(int(nbytes), 256)
    key = (like.device.index, torch.cuda.current_stream(like.device).cuda_stream)
    buf = _WS_CACHE.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(nbytes, dtype=torch.uint8, device=like.device)
        _WS_CACHE[key] = buf
    return buf


# Zero-initialised scratch for kernels that keep a self-resetting ticket counter in it (cp_update with a fused Gram).
_ZWS_CACHE: dict = {}


def _zero_workspace(nbytes: int, like: torch.Tensor) -> torch.Tensor:
    nbytes = max(int(nbytes), 256)
    key = (like.device.index, torch.cuda.current_stream(like.device).cuda_stream)
    buf = _ZWS_CACHE.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.zeros(nbytes, dtype=torch.uint8, device=like.device)
        _ZWS_CACHE[key] = buf
    return buf


def release_workspaces() -> None:
    """Drop the cached scratch buffers (they are re-created on demand)."""
    _WS_CACHE.clear()
    _ZWS_CACHE.clear()


def _prod(xs: Iterable[int]) -> int:
    p = 1
    for x in xs:
        p *= int(x)
    return p


# --------------------------------------------------------------------------- unfold / fold
def unfold(tensor: torch.Tensor, mode: int, contiguous: bool = False) -> torch.Tensor:
    """Mode-`mode` unfolding (tensorly/base.py:39-53), bit-exact.

    Mode 0 is a free view of a contiguous tensor, exactly as in the reference; every other
    mode is produced by the permuting-copy kernel as a C-contiguous matrix.
    """
    _check_tensor(tensor, "tensor")
    ndim = tensor.dim()
    if ndim < 1 or ndim > _lib.MAX_NDIM:
        raise ValueError(f"unfold supports 1..{_lib.MAX_NDIM}-way tensors, got {ndim}")
    if not -ndim <= mode < ndim:
        raise ValueError(f"mode {mode} out of range for a {ndim}-way tensor")
    mode %= ndim
    x = tensor if tensor.is_contiguous() else tensor.contiguous()
    rows = x.shape[mode]
    if mode == 0 and not contiguous:
        return x.reshape(rows, -1)
    out = torch.empty((rows, x.numel() // max(rows, 1)), dtype=x.dtype, device=x.device)
    if x.numel() == 0:
        return out
    lib = _lib.load()
    with _Device(x):
        st = lib.tlb200_unfold(x.data_ptr(), _lib.i64_array(x.shape), ndim, mode, _DTYPES[x.dtype], out.data_ptr(),
                               _stream(x))
    _lib.check(st, "unfold")
    return out


def fold(unfolded_tensor: torch.Tensor, mode: int, shape: Sequence[int]) -> torch.Tensor:
    """Inverse of unfold (tensorly/base.py:56-79); returns a C-contiguous tensor."""
    _check_tensor(unfolded_tensor, "unfolded_tensor")
    shape = tuple(int(s) for s in shape)
    ndim = len(shape)
    if ndim < 1 or ndim > _lib.MAX_NDIM:
        raise ValueError(f"fold supports 1..{_lib.MAX_NDIM}-way shapes, got {ndim}")
    if not -ndim <= mode < ndim:
        raise ValueError(f"mode {mode} out of range for shape {shape}")
    mode %= ndim
    if unfolded_tensor.dim() != 2 or unfolded_tensor.shape[0] != shape[mode] or unfolded_tensor.numel() != _prod(shape):
        raise ValueError(f"cannot fold a matrix of shape {tuple(unfolded_tensor.shape)} into {shape} along mode {mode}")
    u = unfolded_tensor if unfolded_tensor.is_contiguous() else unfolded_tensor.contiguous()
    out = torch.empty(shape, dtype=u.dtype, device=u.device)
    if u.numel() == 0:
        return out
    lib = _lib.load()
    with _Device(u):
        st = lib.tlb200_fold(u.data_ptr(), _lib.i64_array(shape), ndim, mode, _DTYPES[u.dtype], out.data_ptr(), _stream(u))
    _lib.check(st, "fold")
    return out


# --------------------------------------------------------------------------- khatri_rao
def khatri_rao(matrices, weights=None, skip_matrix=None, mask=None) -> torch.Tensor:
    """Khatri-Rao product (tensorly/tenalg/core_tenalg/_khatri_rao.py:9-109), bit-exact:
    first matrix slowest, left fold of multiplies, weights folded into the first matrix,
    optional row mask.  One remaining matrix is returned as is (weights ignored), 1-D
    inputs are treated as single columns with a warning, shape errors raise ValueError.
    """
    matrices = list(matrices)
    if skip_matrix is not None:
        matrices = [matrices[i] for i in range(len(matrices)) if i != skip_matrix]
    if len(matrices) == 0:
        raise ValueError("khatri_rao needs at least one matrix")
    if len(matrices) == 1:
        return matrices[0]
    first = _check_tensor(matrices[0], "matrices[0]")
    if first.dim() == 2:
        n_columns = first.shape[1]
    else:
        n_columns = 1
        matrices = [_check_tensor(m, f"matrices[{i}]", first).reshape(-1, 1) for i, m in enumerate(matrices)]
        warnings.warn("Khatri-rao of a series of vectors instead of matrices. "
                      "Considering each as a matrix with 1 column.")
    for i, m in enumerate(matrices):
        _check_tensor(m, f"matrices[{i}]", first)
        if m.dim() != 2:
            raise ValueError("All the matrices must have exactly 2 dimensions!"
                             f"Matrix {i} has dimension {m.dim()} != 2.")
        if m.shape[1] != n_columns:
            raise ValueError("All matrices must have same number of columns!"
                             f"Matrix {i} has {m.shape[1]} columns != {n_columns}.")
    if len(matrices) > _lib.MAX_NDIM:
        raise ValueError(f"khatri_rao supports at most {_lib.MAX_NDIM} matrices")
    rows = [m.shape[0] for m in matrices]
    total = _prod(rows)
    w = None
    if weights is not None:
        w = torch.as_tensor(weights, dtype=first.dtype, device=first.device).reshape(-1).contiguous()
        if w.numel() != n_columns:
            raise ValueError(f"weights has {w.numel()} entries but the matrices have {n_columns} columns")
    mk = None
    if mask is not None:
        mk = torch.as_tensor(mask, device=first.device).to(first.dtype).reshape(-1).contiguous()
        if mk.numel() != total:
            raise ValueError(f"mask has {mk.numel()} entries but the product has {total} rows")
    out = torch.empty((total, n_columns), dtype=first.dtype, device=first.device)
    if out.numel() == 0:
        return out
    lib = _lib.load()
    with _Device(first):
        st = lib.tlb200_khatri_rao(
            _lib.ptr_array(m.data_ptr() for m in matrices), _lib.i64_array(rows),
            _lib.i64_array(m.stride(0) for m in matrices), _lib.i64_array(m.stride(1) for m in matrices),
            len(matrices), n_columns, w.data_ptr() if w is not None else None,
            mk.data_ptr() if mk is not None else None, _DTYPES[first.dtype], out.data_ptr(), n_columns, _stream(first))
    _lib.check(st, "khatri_rao")
    return out


# --------------------------------------------------------------------------- MTTKRP
def mttkrp_plan(shape: Sequence[int], mode: int, rank: int, dtype=torch.float32, path: str | None = None):
    """The (A, J, B) streaming plan the MTTKRP kernel would use (host-only, no GPU work)."""
    lib = _lib.load()
    plan = _lib.MttkrpPlan()
    st = lib.tlb200_mttkrp_plan(_lib.i64_array(shape), len(shape), mode, rank, _DTYPES[dtype],
                                _lib.PATHS[path or _path], plan)
    _lib.check(st, "mttkrp_plan")
    return plan


def _check_out(out, rows, rank, like):
    _check_tensor(out, "out", like)
    if tuple(out.shape) != (rows, rank) or not out.is_contiguous():
        raise ValueError(f"out must be a contiguous ({rows}, {rank}) tensor")
    return out


def unfolding_dot_khatri_rao(tensor: torch.Tensor, cp_tensor, mode: int) -> torch.Tensor:
    """MTTKRP: dot(unfold(tensor, mode), khatri_rao(factors, weights, skip_matrix=mode))
    (tensorly/tenalg/core_tenalg/mttkrp.py:9-50) without materialising either operand.

    `cp_tensor` is any 2-iterable `(weights | None, factors)`; factors may have arbitrary
    strides.  Returns a new (tensor.shape[mode], rank) tensor.
    """
    return mttkrp(tensor, cp_tensor, mode)


class PartialMttkrp:
    """Unsummed split-K partials of an MTTKRP (tlb200_partials_t) plus the workspace that holds them.  Valid until the
    next tlb200 call that uses the per-stream workspace; consumed by cp_update_fused."""

    def __init__(self, info, ws, dtype, device):
        self.info, self.ws, self.dtype, self.device = info, ws, dtype, device
        self.shape = (int(info.rows), int(info.rank))


def plain_partials(m: torch.Tensor) -> "PartialMttkrp":
    """A summed MTTKRP (rows x rank, unit column stride) dressed as a one-split PartialMttkrp for cp_update_fused."""
    _check_tensor(m, "m")
    if m.dim() != 2 or m.stride(1) != 1:
        raise ValueError("plain_partials expects a row-major matrix")
    info = _lib.Partials()
    info.data, info.splits, info.split_stride = m.data_ptr(), 1, 0
    info.ld, info.rows, info.rank = m.stride(0), m.shape[0], m.shape[1]
    return PartialMttkrp(info, m, m.dtype, m.device)


def mttkrp_partials(tensor: torch.Tensor, cp_tensor, mode: int) -> "PartialMttkrp":
    """unfolding_dot_khatri_rao without its final split-K reduction launch (the solve sums the partials while it loads
    them).  Raises NotImplementedError where the MTTKRP needs several passes (rank > 64 on the tensor-core path)."""
    return mttkrp(tensor, cp_tensor, mode, _partials=True)


def mttkrp(tensor: torch.Tensor, cp_tensor, mode: int, out: torch.Tensor | None = None, _partials: bool = False):
    """unfolding_dot_khatri_rao with an optional preallocated contiguous result (rows, rank) — the reference
    signature has no such argument, the drivers use it to place the result inside a packed all-reduce buffer."""
    _check_tensor(tensor, "tensor")
    weights, factors = cp_tensor
    factors = list(factors)
    ndim = tensor.dim()
    if ndim < 2 or ndim > _lib.MAX_NDIM:
        raise ValueError(f"unfolding_dot_khatri_rao supports 2..{_lib.MAX_NDIM}-way tensors, got {ndim}")
    if len(factors) != ndim:
        raise ValueError(f"got {len(factors)} factors for a {ndim}-way tensor")
    if not -ndim <= mode < ndim:
        raise ValueError(f"mode {mode} out of range for a {ndim}-way tensor")
    mode %= ndim
    rank = None
    for i, f in enumerate(factors):
        if i == mode:
            continue
        _check_tensor(f, f"factors[{i}]", tensor)
        if f.dim() != 2:
            raise ValueError(f"factors[{i}] must be a matrix, got {f.dim()} dimensions")
        if f.shape[0] != tensor.shape[i]:
            raise ValueError(f"factors[{i}] has {f.shape[0]} rows but the tensor has extent {tensor.shape[i]} in mode {i}")
        if rank is None:
            rank = f.shape[1]
        elif f.shape[1] != rank:
            raise ValueError("All matrices must have same number of columns!"
                             f"Matrix {i} has {f.shape[1]} columns != {rank}.")
    if rank is None or rank < 1:
        raise ValueError("rank must be >= 1")
    w = None
    if weights is not None:
        w = torch.as_tensor(weights, dtype=tensor.dtype, device=tensor.device).reshape(-1).contiguous()
        if w.numel() != rank:
            raise ValueError(f"weights has {w.numel()} entries but the factors have {rank} columns")
    x = tensor if tensor.is_contiguous() else tensor.contiguous()
    if not _partials:
        out = (torch.empty((x.shape[mode], rank), dtype=x.dtype, device=x.device) if out is None
               else _check_out(out, x.shape[mode], rank, x))
        if x.numel() == 0:
            return out.zero_()
    elif x.numel() == 0:
        raise NotImplementedError("mttkrp_partials of an empty tensor")
    lib = _lib.load()
    dt = _DTYPES[x.dtype]
    path = _lib.PATHS[_path]
    shape = _lib.i64_array(x.shape)
    nbytes = lib.tlb200_mttkrp_workspace_bytes(shape, ndim, mode, rank, dt, path)
    if nbytes == 0:
        raise ValueError(f"unfolding_dot_khatri_rao: unsupported problem shape={tuple(x.shape)} mode={mode} rank={rank}")
    ws = _workspace(nbytes, x)
    ptrs = [0 if i == mode else f.data_ptr() for i, f in enumerate(factors)]
    rs = [0 if i == mode else f.stride(0) for i, f in enumerate(factors)]
    cs = [0 if i == mode else f.stride(1) for i, f in enumerate(factors)]
    if _partials:
        info = _lib.Partials()
        with _Device(x):
            st = lib.tlb200_mttkrp_partials(x.data_ptr(), shape, ndim, mode, _lib.ptr_array(ptrs), _lib.i64_array(rs),
                                            _lib.i64_array(cs), rank, w.data_ptr() if w is not None else None, dt,
                                            ws.data_ptr(), ws.numel(), path, info, _stream(x))
        if st == _lib.TLB200_EUNSUPPORTED:
            raise NotImplementedError("mttkrp_partials: this problem runs in several passes")
        _lib.check(st, "mttkrp_partials")
        return PartialMttkrp(info, ws, x.dtype, x.device)
    with _Device(x):
        st = lib.tlb200_mttkrp(x.data_ptr(), shape, ndim, mode, _lib.ptr_array(ptrs), _lib.i64_array(rs),
                               _lib.i64_array(cs), rank, w.data_ptr() if w is not None else None, dt, out.data_ptr(),
                               rank, ws.data_ptr(), ws.numel(), path, _stream(x))
    _lib.check(st, "unfolding_dot_khatri_rao")
    return out


def mttkrp_from_ttm_partials(contracted: torch.Tensor, cp_tensor, mode: int) -> "PartialMttkrp":
    """mttkrp_from_ttm without its final reduction launch (see mttkrp_partials)."""
    return mttkrp_from_ttm(contracted, cp_tensor, mode, _partials=True)


def mttkrp_from_ttm(contracted: torch.Tensor, cp_tensor, mode: int, out: torch.Tensor | None = None, _partials: bool = False):
    """MTTKRP of mode `mode` < N-1 from T = mode_dot(tensor, factors[N-1], N-1, transpose=True)
    (shape I_0 x .. x I_{N-2} x rank): equals unfolding_dot_khatri_rao(tensor, cp_tensor, mode) while
    factors[N-1] is unchanged, reading T (rank / I_{N-1} of the tensor) instead of the tensor.
    `cp_tensor` = (weights | None, factors) with all N factors (entries `mode` and N-1 are not read)."""
    _check_tensor(contracted, "contracted")
    weights, factors = cp_tensor
    factors = list(factors)
    nlead = contracted.dim() - 1
    if nlead < 2 or nlead + 1 > _lib.MAX_NDIM:
        raise ValueError(f"mttkrp_from_ttm needs a 3..{_lib.MAX_NDIM}-way problem, got {nlead + 1}")
    if len(factors) != nlead + 1:
        raise ValueError(f"got {len(factors)} factors for a {nlead + 1}-way tensor")
    if not 0 <= mode < nlead:
        raise ValueError(f"mode {mode} must be one of the first {nlead} modes")
    rank = contracted.shape[-1]
    for i in range(nlead):
        if i == mode:
            continue
        f = factors[i]
        _check_tensor(f, f"factors[{i}]", contracted)
        if f.dim() != 2 or f.shape[0] != contracted.shape[i] or f.shape[1] != rank:
            raise ValueError(f"factors[{i}] must be ({contracted.shape[i]}, {rank}), got {tuple(f.shape)}")
    w = None
    if weights is not None:
        w = torch.as_tensor(weights, dtype=contracted.dtype, device=contracted.device).reshape(-1).contiguous()
        if w.numel() != rank:
            raise ValueError(f"weights has {w.numel()} entries but the factors have {rank} columns")
    t = contracted if contracted.is_contiguous() else contracted.contiguous()
    if not _partials:
        out = (torch.empty((t.shape[mode], rank), dtype=t.dtype, device=t.device) if out is None
               else _check_out(out, t.shape[mode], rank, t))
        if t.numel() == 0:
            return out.zero_()
    elif t.numel() == 0:
        raise NotImplementedError("mttkrp_from_ttm_partials of an empty tensor")
    lib = _lib.load()
    dt = _DTYPES[t.dtype]
    lead = _lib.i64_array(t.shape[:nlead])
    nbytes = lib.tlb200_mttkrp_from_ttm_workspace_bytes(lead, nlead, mode, rank, dt)
    if nbytes == 0:
        raise ValueError(f"mttkrp_from_ttm: unsupported problem shape={tuple(t.shape)} mode={mode}")
    ws = _workspace(nbytes, t)
    ptrs = [0 if i == mode else factors[i].data_ptr() for i in range(nlead)]
    rs = [0 if i == mode else factors[i].stride(0) for i in range(nlead)]
    cs = [0 if i == mode else factors[i].stride(1) for i in range(nlead)]
    if _partials:
        info = _lib.Partials()
        with _Device(t):
            st = lib.tlb200_mttkrp_from_ttm_partials(t.data_ptr(), lead, nlead, mode, _lib.ptr_array(ptrs), _lib.i64_array(rs),
                                                     _lib.i64_array(cs), rank, w.data_ptr() if w is not None else None, dt,
                                                     ws.data_ptr(), ws.numel(), info, _stream(t))
        _lib.check(st, "mttkrp_from_ttm_partials")
        return PartialMttkrp(info, ws, t.dtype, t.device)
    with _Device(t):
        st = lib.tlb200_mttkrp_from_ttm(t.data_ptr(), lead, nlead, mode, _lib.ptr_array(ptrs), _lib.i64_array(rs),
                                        _lib.i64_array(cs), rank, w.data_ptr() if w is not None else None, dt,
                                        out.data_ptr(), rank, ws.data_ptr(), ws.numel(), _stream(t))
    _lib.check(st, "mttkrp_from_ttm")
    return out


# --------------------------------------------------------------------------- TTM
def mode_dot(tensor: torch.Tensor, matrix_or_vector: torch.Tensor, mode: int, transpose: bool = False) -> torch.Tensor:
    """n-mode product with a matrix or a vector (n_mode_product.py:5-76).

    matrix: (J, I_mode) [or (I_mode, J) with transpose=True] -> mode extent becomes J;
    vector: (I_mode,) -> the mode is dropped.  ValueError on extent mismatch.
    """
    _check_tensor(tensor, "tensor")
    _check_tensor(matrix_or_vector, "matrix_or_vector", tensor)
    ndim = tensor.dim()
    if ndim < 1 or ndim > _lib.MAX_NDIM:
        raise ValueError(f"mode_dot supports 1..{_lib.MAX_NDIM}-way tensors, got {ndim}")
    if not -ndim <= mode < ndim:
        raise ValueError(f"mode {mode} out of range for a {ndim}-way tensor")
    mode %= ndim
    m = matrix_or_vector
    new_shape = list(tensor.shape)
    if m.dim() == 2:
        dim = 0 if transpose else 1
        if m.shape[dim] != tensor.shape[mode]:
            raise ValueError(
                f"shapes {tuple(tensor.shape)} and {tuple(m.shape)} not aligned in mode-{mode} multiplication: "
                f"{tensor.shape[mode]} (mode {mode}) != {m.shape[dim]} (dim 1 of matrix)")
        rows_out = m.shape[0 if not transpose else 1]
        rs, cs = (m.stride(0), m.stride(1)) if not transpose else (m.stride(1), m.stride(0))
        new_shape[mode] = rows_out
        final_shape = tuple(new_shape)
    elif m.dim() == 1:
        if m.shape[0] != tensor.shape[mode]:
            raise ValueError(
                f"shapes {tuple(tensor.shape)} and {tuple(m.shape)} not aligned for mode-{mode} multiplication: "
                f"{tensor.shape[mode]} (mode {mode}) != {m.shape[0]} (vector size)")
        rows_out, rs, cs = 1, 0, m.stride(0)
        new_shape[mode] = 1
        final_shape = tuple(s for i, s in enumerate(tensor.shape) if i != mode) if ndim > 1 else ()
    else:
        raise ValueError("Can only take n_mode_product with a vector or a matrix."
                         f"Provided array of dimension {m.dim()} not in [1, 2].")
    x = tensor if tensor.is_contiguous() else tensor.contiguous()
    out = torch.empty(tuple(new_shape), dtype=x.dtype, device=x.device)
    if out.numel() == 0:
        return out.reshape(final_shape)
    if x.numel() == 0:
        return out.zero_().reshape(final_shape)
    lib = _lib.load()
    shape = _lib.i64_array(x.shape)
    dt, path = _DTYPES[x.dtype], _lib.PATHS[_path]
    ws = _workspace(lib.tlb200_mode_dot_workspace_bytes(shape, ndim, mode, rows_out, dt, path), x)
    with _Device(x):
        st = lib.tlb200_mode_dot(x.data_ptr(), shape, ndim, mode, m.data_ptr(), rows_out, rs, cs, dt, out.data_ptr(),
                                 ws.data_ptr(), ws.numel(), path, _stream(x))
    _lib.check(st, "mode_dot")
    return out.reshape(final_shape)


def multi_mode_dot(tensor: torch.Tensor, matrix_or_vec_list, modes=None, skip=None, transpose: bool = False) -> torch.Tensor:
    """Chain of n-mode products (n_mode_product.py:79-135): pairs sorted by mode, `skip`
    indexes the sorted list, vectors drop their mode.  Distinct modes run as one fused
    C-ABI call (intermediates stay in a workspace); repeated modes fall back to
    successive mode_dot calls.
    """
    _check_tensor(tensor, "tensor")
    matrix_or_vec_list = list(matrix_or_vec_list)
    if modes is None:
        modes = range(len(matrix_or_vec_list))
    modes = [int(m) for m in modes]
    pairs = sorted(zip(matrix_or_vec_list, modes), key=lambda p: p[1])
    pairs = [p for i, p in enumerate(pairs) if not (skip is not None and i == skip)]
    ndim = tensor.dim()
    used = [m for _, m in pairs]
    if len(set(used)) != len(used) or any(not 0 <= m < ndim for m in used):
        res, decrement = tensor, 0
        for mat, mode in pairs:
            res = mode_dot(res, mat, mode - decrement, transpose=transpose)
            if mat.dim() == 1:
                decrement += 1
        return res
    if not pairs:
        return tensor.clone()
    if len(pairs) == 1:
        return mode_dot(tensor, pairs[0][0], pairs[0][1], transpose=transpose)
    rows_out, rss, css, ptrs = [], [], [], []
    new_shape = list(tensor.shape)
    drop = []
    for mat, mode in pairs:
        _check_tensor(mat, "matrix_or_vec_list entry", tensor)
        if mat.dim() == 2:
            dim = 0 if transpose else 1
            if mat.shape[dim] != tensor.shape[mode]:
                raise ValueError(
                    f"shapes {tuple(tensor.shape)} and {tuple(mat.shape)} not aligned in mode-{mode} multiplication: "
                    f"{tensor.shape[mode]} (mode {mode}) != {mat.shape[dim]} (dim 1 of matrix)")
            r = mat.shape[1 if transpose else 0]
            rs, cs = (mat.stride(1), mat.stride(0)) if transpose else (mat.stride(0), mat.stride(1))
        elif mat.dim() == 1:
            if mat.shape[0] != tensor.shape[mode]:
                raise ValueError(
                    f"shapes {tuple(tensor.shape)} and {tuple(mat.shape)} not aligned for mode-{mode} multiplication: "
                    f"{tensor.shape[mode]} (mode {mode}) != {mat.shape[0]} (vector size)")
            r, rs, cs = 1, 0, mat.stride(0)
            drop.append(mode)
        else:
            raise ValueError("Can only take n_mode_product with a vector or a matrix."
                             f"Provided array of dimension {mat.dim()} not in [1, 2].")
        rows_out.append(r); rss.append(rs); css.append(cs); ptrs.append(mat.data_ptr())
        new_shape[mode] = r
    final_shape = tuple(s for i, s in enumerate(new_shape) if i not in drop)
    x = tensor if tensor.is_contiguous() else tensor.contiguous()
    out = torch.empty(tuple(new_shape), dtype=x.dtype, device=x.device)
    if out.numel() == 0 or x.numel() == 0:
        return out.zero_().reshape(final_shape)
    lib = _lib.load()
    dt = _DTYPES[x.dtype]
    path = _lib.PATHS[_path]
    shape = _lib.i64_array(x.shape)
    cmodes = _lib.int_array(used)
    crows = _lib.i64_array(rows_out)
    nbytes = lib.tlb200_multi_mode_dot_workspace_bytes(shape, ndim, cmodes, crows, len(pairs), dt, path)
    ws = _workspace(nbytes, x)
    with _Device(x):
        st = lib.tlb200_multi_mode_dot(x.data_ptr(), shape, ndim, cmodes, _lib.ptr_array(ptrs), crows,
                                       _lib.i64_array(rss), _lib.i64_array(css), len(pairs), dt, out.data_ptr(),
                                       ws.data_ptr(), ws.numel(), path, _stream(x))
    _lib.check(st, "multi_mode_dot")
    return out.reshape(final_shape)


# --------------------------------------------------------------------------- CP-ALS pieces
def sumsq(x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """||x||^2 as a device scalar of x.dtype (double accumulation, deterministic)."""
    _check_tensor(x, "x")
    xc = x if x.is_contiguous() else x.contiguous()
    if out is None:
        out = torch.empty(1, dtype=x.dtype, device=x.device)
    lib = _lib.load()
    dt = _DTYPES[x.dtype]
    ws = _workspace(lib.tlb200_sumsq_workspace_bytes(xc.numel(), dt), x)
    with _Device(x):
        st = lib.tlb200_sumsq(xc.data_ptr(), xc.numel(), dt, out.data_ptr(), ws.data_ptr(), ws.numel(), _stream(x))
    _lib.check(st, "sumsq")
    return out


def tensor_absmax(x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """max |x| as a device float (fp32 tensors; one pass, no host sync) — the range hint of the fp16-split engine."""
    _check_tensor(x, "x")
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise NotImplementedError("tensor_absmax takes contiguous float32 tensors")
    if out is None:
        out = torch.empty(1, dtype=torch.float32, device=x.device)
    lib = _lib.load()
    with _Device(x):
        st = lib.tlb200_tensor_absmax(x.data_ptr(), x.numel(), _DTYPES[x.dtype], out.data_ptr(), _stream(x))
    _lib.check(st, "tensor_absmax")
    return out


class RangeHint:
    """Registers max |x| of a tensor for as long as this object lives: MTTKRP / mode_dot calls on that tensor (same
    base pointer) then run on the fp16-split tensor-core engine instead of 3xTF32 (include/tlb200.h,
    tlb200_hint_tensor_absmax).  The owner promises not to change the tensor while the hint is registered — the
    ALS drivers hold one for their (constant) input tensor.  `TLB200_DISABLE_HF=1` turns the engine off."""

    _live: dict = {}            # base pointer -> [(id of hint, absmax tensor), ...] in registration order

    def __init__(self, x: torch.Tensor, hold: bool = True):
        self.ptr = None
        self.absmax = tensor_absmax(x)
        # hold=True keeps the tensor (and therefore the pointer's meaning) alive; hold=False withdraws the hint the
        # moment the tensor object is collected instead
        self._x = x if hold else None
        _lib.check(_lib.load().tlb200_hint_tensor_absmax(x.data_ptr(), self.absmax.data_ptr()), "hint_tensor_absmax")
        self.ptr = x.data_ptr()
        RangeHint._live.setdefault(self.ptr, []).append((id(self), self.absmax))
        if not hold:
            weakref.finalize(x, RangeHint._withdraw, self.ptr, id(self))

    @staticmethod
    def _withdraw(ptr, owner):
        """Several owners may hold a hint for the same tensor (two drivers on one input): the registration survives
        until the last of them lets go, and always points at a live scalar."""
        try:
            entries = RangeHint._live.get(ptr)
            if not entries:
                return
            entries[:] = [e for e in entries if e[0] != owner]
            lib = _lib.load()
            if entries:
                lib.tlb200_hint_tensor_absmax(ptr, entries[-1][1].data_ptr())
            else:
                del RangeHint._live[ptr]
                lib.tlb200_hint_tensor_absmax(ptr, None)
        except Exception:               # interpreter shutdown
            pass

    @staticmethod
    def applies(x: torch.Tensor, rank: int) -> bool:
        """fp32 CUDA tensors (measured: +4..7 % streamed bytes/s at rank 64, +5 % sustained at rank 32, under the power cap)."""
        lo = int(os.environ.get("TLB200_HF_MIN_RANK", "1"))
        return (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and rank >= lo
                and os.environ.get("TLB200_DISABLE_HF", "0") in ("", "0"))

    def close(self) -> None:
        if self.ptr is not None:
            try:
                RangeHint._withdraw(self.ptr, id(self))     # (a later hint for the same tensor stays)
            except Exception:           # interpreter shutdown: the class itself may be gone
                pass
            self.ptr = None
            self._x = None

    def __del__(self):
        self.close()


def gram(f: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """F^T F (rank x rank), deterministic."""
    _check_tensor(f, "f")
    if f.dim() != 2:
        raise ValueError("gram expects a matrix")
    rows, rank = f.shape
    if out is None:
        out = torch.empty((rank, rank), dtype=f.dtype, device=f.device)
    lib = _lib.load()
    dt = _DTYPES[f.dtype]
    ws = _workspace(lib.tlb200_gram_workspace_bytes(rows, rank, dt), f)
    with _Device(f):
        st = lib.tlb200_gram(f.data_ptr(), rows, rank, f.stride(0), f.stride(1), dt, out.data_ptr(), ws.data_ptr(),
                             ws.numel(), _stream(f))
    _lib.check(st, "gram")
    return out


def _gram_ptrs(grams, skip):
    return _lib.ptr_array(0 if (g is None or i == skip) else g.data_ptr() for i, g in enumerate(grams))


def cp_update(grams, mode: int, weights, mttkrp: torch.Tensor, l2_reg: float = 0.0,
              out: torch.Tensor | None = None, gram_out: torch.Tensor | None = None) -> torch.Tensor:
    """ALS factor update: solve(V^T, M^T)^T with V = (w w^T) o prod_{i != mode} G_i + l2 I
    (tensorly/decomposition/_cp.py:411-428).  With `gram_out` (rank x rank, may be grams[mode]) the same
    launch also writes the Gram matrix of the updated factor."""
    _check_tensor(mttkrp, "mttkrp")
    rows, rank = mttkrp.shape
    if mttkrp.stride(1) != 1:
        mttkrp = mttkrp.contiguous()
    if out is None:
        out = torch.empty((rows, rank), dtype=mttkrp.dtype, device=mttkrp.device)
    lib = _lib.load()
    if gram_out is not None and rows > 0:
        _check_tensor(gram_out, "gram_out", mttkrp)
        if tuple(gram_out.shape) != (rank, rank) or not gram_out.is_contiguous():
            raise ValueError("gram_out must be a contiguous (rank, rank) tensor")
        dt = _DTYPES[mttkrp.dtype]
        with _Device(mttkrp):
            nbytes = lib.tlb200_cp_update_gram_workspace_bytes(rows, rank, dt)
            ws = _zero_workspace(nbytes, mttkrp)
            st = lib.tlb200_cp_update_gram(_gram_ptrs(grams, mode), len(grams), mode, rank,
                                           weights.data_ptr() if weights is not None else None, float(l2_reg or 0.0),
                                           mttkrp.data_ptr(), mttkrp.stride(0), rows, dt, out.data_ptr(), out.stride(0),
                                           gram_out.data_ptr(), ws.data_ptr(), ws.numel(), _stream(mttkrp))
        _lib.check(st, "cp_update_gram")
        return out
    with _Device(mttkrp):
        st = lib.tlb200_cp_update(_gram_ptrs(grams, mode), len(grams), mode, rank,
                                  weights.data_ptr() if weights is not None else None, float(l2_reg or 0.0),
                                  mttkrp.data_ptr(), mttkrp.stride(0), rows, _DTYPES[mttkrp.dtype], out.data_ptr(),
                                  out.stride(0), _stream(mttkrp))
    _lib.check(st, "cp_update")
    return out


def nncp_update(grams, mode: int, weights, mttkrp: torch.Tensor, factor: torch.Tensor, eps: float) -> torch.Tensor:
    """In-place multiplicative update of `factor` (tensorly/decomposition/_nn_cp.py:131-136)."""
    _check_tensor(mttkrp, "mttkrp")
    _check_tensor(factor, "factor", mttkrp)
    rows, rank = factor.shape
    if factor.stride(1) != 1 or mttkrp.stride(1) != 1:
        raise ValueError("nncp_update needs row-major factor and mttkrp")
    lib = _lib.load()
    with _Device(factor):
        st = lib.tlb200_nncp_update(_gram_ptrs(grams, mode), len(grams), mode, rank,
                                    weights.data_ptr() if weights is not None else None, mttkrp.data_ptr(),
                                    mttkrp.stride(0), factor.data_ptr(), factor.stride(0), rows, float(eps),
                                    _DTYPES[factor.dtype], _stream(factor))
    _lib.check(st, "nncp_update")
    return factor


def cp_error(grams, weights, mttkrp_last: torch.Tensor, factor_last: torch.Tensor, norm_x2: torch.Tensor,
             out: torch.Tensor | None = None) -> torch.Tensor:
    """Relative reconstruction error from the last mode's MTTKRP (_cp.py:217-225):
    out = [rel_error, iprod, ||cp||^2] as device scalars (no host sync)."""
    _check_tensor(mttkrp_last, "mttkrp_last")
    rows, rank = mttkrp_last.shape
    if out is None:
        out = torch.empty(3, dtype=mttkrp_last.dtype, device=mttkrp_last.device)
    lib = _lib.load()
    with _Device(mttkrp_last):
        st = lib.tlb200_cp_error(_gram_ptrs(grams, -1), len(grams), rank,
                                 weights.data_ptr() if weights is not None else None, mttkrp_last.data_ptr(),
                                 mttkrp_last.stride(0), factor_last.data_ptr(), factor_last.stride(0),
                                 factor_last.stride(1), rows, norm_x2.data_ptr(), _DTYPES[mttkrp_last.dtype],
                                 out.data_ptr(), _stream(mttkrp_last))
    _lib.check(st, "cp_error")
    return out


# --------------------------------------------------------------------------- reconstruction
def _check_cp(cp_tensor, what: str):
    """(weights | None, factors) -> (w tensor | None, factors, shape, rank) with the reference's validation
    (tensorly/cp_tensor.py:163-214): matrices with one common column count, weights of that length."""
    weights, factors = cp_tensor
    factors = list(factors)
    if len(factors) < 1:
        raise ValueError(f"{what}: a CP tensor needs at least one factor")
    first = _check_tensor(factors[0], "factors[0]")
    for i, f in enumerate(factors):
        _check_tensor(f, f"factors[{i}]", first)
        if f.dim() != 2:
            raise ValueError("A CP tensor should be composed of a list of matrices,"
                             f"but factor {i} has {f.dim()} dimensions.")
        if f.shape[1] != first.shape[1]:
            raise ValueError("All the factors of a CP tensor should have the same number of column."
                             f"However, factors[0].shape[1]={first.shape[1]} but factors[{i}].shape[1]={f.shape[1]}.")
    rank = first.shape[1]
    w = None
    if weights is not None:
        w = torch.as_tensor(weights, dtype=first.dtype, device=first.device).reshape(-1).contiguous()
        if w.numel() != rank:
            raise ValueError(f"Given factors for a rank-{rank} CP tensor but len(weights)={w.numel()}.")
    return w, factors, tuple(f.shape[0] for f in factors), rank


def cp_to_tensor(cp_tensor, mask=None, out: torch.Tensor | None = None) -> torch.Tensor:
    """Full tensor of a CP decomposition (tensorly/cp_tensor.py:433-485): sum_r w_r a_r o b_r o c_r ..., written
    once, without the reference's Khatri-Rao matrix.  `mask` (an array with as many entries as the tensor)
    multiplies the result element-wise — the evident intent of the reference's `mask` branch, whose
    `khatri_rao(..., mask=mask)` only accepts a column of prod(shape) entries."""
    w, factors, shape, rank = _check_cp(cp_tensor, "cp_to_tensor")
    first = factors[0]
    vector = len(shape) == 1
    if vector:                 # reference: sum(weights * factors[0], axis=1); here the 2-way case with a row of ones
        factors = factors + [torch.ones((1, rank), dtype=first.dtype, device=first.device)]
        shape = shape + (1,)
    if len(shape) > _lib.MAX_NDIM:
        raise ValueError(f"cp_to_tensor supports at most {_lib.MAX_NDIM} modes")
    mk = None
    if mask is not None:
        mk = torch.as_tensor(mask, device=first.device).to(first.dtype).contiguous()
        if mk.numel() != _prod(shape):
            raise ValueError(f"mask has {mk.numel()} entries but the tensor has {_prod(shape)}")
    if out is None:
        out = torch.empty(shape, dtype=first.dtype, device=first.device)
    else:
        _check_tensor(out, "out", first)
        if tuple(out.shape) != shape or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous tensor of shape {shape}")
    if out.numel() > 0:
        lib = _lib.load()
        with _Device(first):
            st = lib.tlb200_cp_to_tensor(_lib.ptr_array(f.data_ptr() for f in factors), _lib.i64_array(shape),
                                         _lib.i64_array(f.stride(0) for f in factors),
                                         _lib.i64_array(f.stride(1) for f in factors), len(shape), rank,
                                         w.data_ptr() if w is not None else None,
                                         mk.data_ptr() if mk is not None else None, _DTYPES[first.dtype],
                                         out.data_ptr(), _stream(first))
        _lib.check(st, "cp_to_tensor")
    return out.reshape(shape[0]) if vector else out


def cp_impute(tensor: torch.Tensor, mask: torch.Tensor, cp_tensor, out: torch.Tensor | None = None,
              stats: torch.Tensor | None = None):
    """Masked-ALS imputation step (tensorly/decomposition/_cp.py:195-207) in one pass:
    out = tensor * mask + cp_to_tensor(cp_tensor) * (1 - mask) (`out` may be `tensor` itself), and
    stats = [||out - rec|| / ||out||, ||out||^2, ||out - rec||^2] as device scalars.  Returns (out, stats)."""
    _check_tensor(tensor, "tensor")
    w, factors, shape, rank = _check_cp(cp_tensor, "cp_impute")
    _check_tensor(factors[0], "factors[0]", tensor)
    if tuple(tensor.shape) != shape:
        raise ValueError(f"tensor has shape {tuple(tensor.shape)} but the factors describe {shape}")
    if tensor.dim() < 2 or tensor.dim() > _lib.MAX_NDIM:
        raise ValueError(f"cp_impute supports 2..{_lib.MAX_NDIM}-way tensors")
    _check_tensor(mask, "mask", tensor)
    if tuple(mask.shape) != shape:
        raise ValueError(f"mask has shape {tuple(mask.shape)} but the tensor has shape {shape}")
    if not tensor.is_contiguous() or not mask.is_contiguous():
        raise ValueError("cp_impute needs C-contiguous tensor and mask")
    if out is None:
        out = torch.empty_like(tensor)
    elif tuple(out.shape) != shape or not out.is_contiguous() or out.dtype != tensor.dtype:
        raise ValueError("out must be a contiguous tensor like `tensor`")
    if stats is None:
        stats = torch.empty(3, dtype=tensor.dtype, device=tensor.device)
    lib = _lib.load()
    cshape = _lib.i64_array(shape)
    ws = _workspace(lib.tlb200_cp_impute_workspace_bytes(cshape, len(shape)), tensor)
    with _Device(tensor):
        st = lib.tlb200_cp_impute(tensor.data_ptr(), mask.data_ptr(), _lib.ptr_array(f.data_ptr() for f in factors), cshape,
                                  _lib.i64_array(f.stride(0) for f in factors),
                                  _lib.i64_array(f.stride(1) for f in factors), len(shape), rank,
                                  w.data_ptr() if w is not None else None, _DTYPES[tensor.dtype], out.data_ptr(),
                                  stats.data_ptr(), ws.data_ptr(), ws.numel(), _stream(tensor))
    _lib.check(st, "cp_impute")
    return out, stats


# --------------------------------------------------------------------------- HOOI pieces
def orthonormalize(z: torch.Tensor, out: torch.Tensor | None = None, passes: int = 2) -> torch.Tensor:
    """Orthonormal basis of the column span of a tall (rows, rank <= 64) block: Cholesky-QR with the small Gram
    matrix, its factor and inverse in fp64 (tlb200_orthonormalize); `passes=2` repeats it on the result (orthonormal
    to rounding), `passes=1` leaves an orthogonality defect of about cond(z)^2 * 1e-16."""
    _check_tensor(z, "z")
    if z.dim() != 2:
        raise ValueError("orthonormalize expects a matrix")
    rows, rank = z.shape
    if rows < rank:
        raise ValueError(f"orthonormalize needs rows >= columns, got {tuple(z.shape)}")
    if out is None:
        out = torch.empty((rows, rank), dtype=z.dtype, device=z.device)
    lib = _lib.load()
    nbytes = lib.tlb200_orthonormalize_workspace_bytes(rows, rank)
    if nbytes == 0:
        raise ValueError(f"orthonormalize: unsupported block {tuple(z.shape)} (at most 64 columns)")
    ws = _zero_workspace(nbytes, z)
    with _Device(z):
        st = lib.tlb200_orthonormalize(z.data_ptr(), rows, rank, z.stride(0), z.stride(1), _DTYPES[z.dtype], out.data_ptr(),
                                       out.stride(0), int(passes), ws.data_ptr(), ws.numel(), _stream(z))
    _lib.check(st, "orthonormalize")
    return out


def symeig(a: torch.Tensor):
    """(eigenvalues descending, eigenvectors as columns) of a small symmetric matrix (order <= 64): Jacobi in fp64 in one
    CTA (tlb200_symeig)."""
    _check_tensor(a, "a")
    if a.dim() != 2 or a.shape[0] != a.shape[1]:
        raise ValueError("symeig expects a square matrix")
    n = a.shape[0]
    if n > 64:
        raise ValueError("symeig supports matrices of order <= 64")
    ac = a if a.is_contiguous() else a.contiguous()
    evals = torch.empty(n, dtype=a.dtype, device=a.device)
    evecs = torch.empty((n, n), dtype=a.dtype, device=a.device)
    lib = _lib.load()
    with _Device(a):
        st = lib.tlb200_symeig(ac.data_ptr(), n, ac.stride(0), _DTYPES[a.dtype], evals.data_ptr(), evecs.data_ptr(), n,
                               _stream(a))
    _lib.check(st, "symeig")
    return evals, evecs


# --------------------------------------------------------------------------- HALS
def hals_update(grams, mode: int, weights, mttkrp: torch.Tensor, factor: torch.Tensor, n_iter_max: int = 100,
                tol: float = 1e-8, sparsity_coefficient=None, ridge_coefficient=None, epsilon: float = 0.0,
                iters_out: torch.Tensor | None = None) -> torch.Tensor:
    """In-place HALS update of `factor` (rows x rank) from its MTTKRP: hals_nnls(M^T, V, F^T) with
    V = (w w^T) o prod_{i != mode} G_i (tensorly/decomposition/_nn_cp.py:311-336, solvers/nnls.py:139-173),
    the whole inner iteration in one kernel.  mode < 0: grams[0] is UtU itself."""
    import ctypes
    _check_tensor(mttkrp, "mttkrp")
    _check_tensor(factor, "factor", mttkrp)
    rows, rank = factor.shape
    if tuple(mttkrp.shape) != (rows, rank):
        raise ValueError(f"mttkrp {tuple(mttkrp.shape)} and factor {tuple(factor.shape)} must have the same shape")
    lib = _lib.load()
    ws = _workspace(lib.tlb200_hals_workspace_bytes(rows), factor)
    sp = ctypes.byref(ctypes.c_double(float(sparsity_coefficient))) if sparsity_coefficient is not None else None
    rg = ctypes.byref(ctypes.c_double(float(ridge_coefficient))) if ridge_coefficient is not None else None
    with _Device(factor):
        st = lib.tlb200_hals_update(_gram_ptrs(grams, mode), len(grams), mode, rank,
                                    weights.data_ptr() if weights is not None else None, mttkrp.data_ptr(),
                                    mttkrp.stride(0), mttkrp.stride(1), factor.data_ptr(), factor.stride(0), factor.stride(1),
                                    rows, int(n_iter_max), float(tol), sp, rg, float(epsilon), _DTYPES[factor.dtype],
                                    iters_out.data_ptr() if iters_out is not None else None, ws.data_ptr(), ws.numel(),
                                    _stream(factor))
    if st == _lib.TLB200_EUNSUPPORTED:
        raise NotImplementedError(f"hals: rank {rank} / {rows} rows exceed the single-kernel solver "
                                  "(rank <= 64 in fp32, 32 in fp64)")
    _lib.check(st, "hals_update")
    return factor


def hals_nnls(UtM, UtU, V=None, n_iter_max=500, tol=1e-8, sparsity_coefficient=None, ridge_coefficient=None,
              nonzero_rows=False, exact=False, epsilon=0.0, callback=None) -> torch.Tensor:
    """tensorly.solvers.nnls.hals_nnls (tensorly/solvers/nnls.py:5-175) for a given V (rank x n): returns the new
    V; the input is not modified.  `nonzero_rows`, `callback` and V=None are not part of the kernel."""
    if V is None or nonzero_rows or callback is not None:
        raise NotImplementedError("hals_nnls kernel: pass V, nonzero_rows=False, callback=None")
    _check_tensor(UtM, "UtM")
    _check_tensor(UtU, "UtU", UtM)
    _check_tensor(V, "V", UtM)
    rank, n = UtM.shape
    if tuple(UtU.shape) != (rank, rank) or tuple(V.shape) != (rank, n):
        raise ValueError(f"shapes: UtM {tuple(UtM.shape)}, UtU {tuple(UtU.shape)}, V {tuple(V.shape)}")
    if exact:
        n_iter_max, tol = 50000, 1e-16
    out = V.clone()
    hals_update([UtU.contiguous()], -1, None, UtM.transpose(0, 1), out.transpose(0, 1), n_iter_max, tol,
                sparsity_coefficient, ridge_coefficient, epsilon)
    return out


def subspace_iterate(g: torch.Tensor, u: torch.Tensor, steps: int) -> torch.Tensor:
    """`steps` power steps u <- orth(g u) IN PLACE (fp64; g symmetric n x n, u n x p with p <= 64): two launches per
    step (tlb200_subspace_iterate).  Returns u."""
    _check_tensor(g, "g")
    _check_tensor(u, "u", g)
    if g.dtype != torch.float64:
        raise TypeError("subspace_iterate runs in float64")
    if g.dim() != 2 or g.shape[0] != g.shape[1] or u.dim() != 2 or u.shape[0] != g.shape[0]:
        raise ValueError(f"subspace_iterate: g {tuple(g.shape)} and u {tuple(u.shape)} do not match")
    if g.stride(1) != 1 or u.stride(1) != 1:
        raise ValueError("subspace_iterate needs row-major g and u")
    n, p = u.shape
    lib = _lib.load()
    nbytes = lib.tlb200_subspace_iterate_workspace_bytes(n, p)
    if nbytes == 0 or n < p:
        raise ValueError(f"subspace_iterate: unsupported block {tuple(u.shape)} (at most 64 columns, rows >= columns)")
    ws = _zero_workspace(nbytes, g)
    with _Device(g):
        st = lib.tlb200_subspace_iterate(g.data_ptr(), n, g.stride(0), u.data_ptr(), p, u.stride(0), int(steps), ws.data_ptr(),
                                         ws.numel(), _stream(g))
    _lib.check(st, "subspace_iterate")
    return u


def cp_update_fused(grams, mode: int, weights, partials: "PartialMttkrp", l2_reg: float = 0.0, out: torch.Tensor | None = None,
                    gram_out: torch.Tensor | None = None, m_out: torch.Tensor | None = None,
                    iprod_out: torch.Tensor | None = None, norm_x2: torch.Tensor | None = None,
                    err_out: torch.Tensor | None = None) -> torch.Tensor:
    """cp_update whose right-hand sides are the unsummed partials of mttkrp_partials / mttkrp_from_ttm_partials: one
    launch sums them (split order), solves, forms the Gram matrix of the new factor and — on request — writes the
    summed MTTKRP (`m_out`) and <M, F_new> (`iprod_out`, a device scalar for cp_error_iprod).  With `err_out` (and
    `norm_x2`, `iprod_out`) — for the LAST mode — the same launch finishes the fast reconstruction error like cp_error."""
    rows, rank = partials.shape
    dtype, device = partials.dtype, partials.device
    if out is None:
        out = torch.empty((rows, rank), dtype=dtype, device=device)
    if gram_out is None:
        gram_out = torch.empty((rank, rank), dtype=dtype, device=device)
    if tuple(gram_out.shape) != (rank, rank) or not gram_out.is_contiguous():
        raise ValueError("gram_out must be a contiguous (rank, rank) tensor")
    lib = _lib.load()
    dt = _DTYPES[dtype]
    with _Device(out):
        ws = _zero_workspace(lib.tlb200_cp_update_gram_workspace_bytes(rows, rank, dt), out)
        st = lib.tlb200_cp_update_fused(_gram_ptrs(grams, mode), len(grams), mode, rank,
                                        weights.data_ptr() if weights is not None else None, float(l2_reg or 0.0),
                                        ctypes_byref(partials.info), dt, out.data_ptr(), out.stride(0), gram_out.data_ptr(),
                                        m_out.data_ptr() if m_out is not None else None,
                                        m_out.stride(0) if m_out is not None else 0,
                                        iprod_out.data_ptr() if iprod_out is not None else None,
                                        norm_x2.data_ptr() if norm_x2 is not None else None,
                                        err_out.data_ptr() if err_out is not None else None, ws.data_ptr(), ws.numel(),
                                        _stream(out))
    if st == _lib.TLB200_EUNSUPPORTED:
        raise NotImplementedError("cp_update_fused: <M, F> is only formed on the register-LU path (rank <= 64 fp32 / 32 fp64)")
    _lib.check(st, "cp_update_fused")
    return out


def ctypes_byref(obj):
    import ctypes
    return ctypes.byref(obj)


def cp_error_iprod(grams, weights, iprod: torch.Tensor, norm_x2: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """cp_error from a device-resident <M_last, F_last> (cp_update_fused's iprod_out)."""
    _check_tensor(iprod, "iprod")
    rank = next(g for g in grams if g is not None).shape[0]
    if out is None:
        out = torch.empty(3, dtype=iprod.dtype, device=iprod.device)
    lib = _lib.load()
    with _Device(iprod):
        st = lib.tlb200_cp_error_iprod(_gram_ptrs(grams, -1), len(grams), rank,
                                       weights.data_ptr() if weights is not None else None, iprod.data_ptr(),
                                       norm_x2.data_ptr(), _DTYPES[iprod.dtype], out.data_ptr(), _stream(iprod))
    _lib.check(st, "cp_error_iprod")
    return out

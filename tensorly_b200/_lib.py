"""ctypes binding of libtlb200.so — the C ABI declared in include/tlb200.h.

There is no fallback: if the shared library is missing or fails to load, importing the
compute entry points raises (the product path is CUDA-only).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libtlb200.so")

TLB200_OK = 0
TLB200_EINVAL = -1
TLB200_EWORKSPACE = -2
TLB200_ECUDA = -3
TLB200_EUNSUPPORTED = -4

F32, F64 = 0, 1
PATH_AUTO, PATH_SIMT, PATH_TCGEN05 = 0, 1, 2
PATHS = {"auto": PATH_AUTO, "simt": PATH_SIMT, "tcgen05": PATH_TCGEN05}
MAX_NDIM = 8


class MttkrpPlan(ctypes.Structure):
    """Mirror of tlb200_mttkrp_plan_t."""

    _fields_ = [
        ("A", c_int64), ("J", c_int64), ("B", c_int64),
        ("sa", c_int64), ("sj", c_int64), ("sb", c_int64),
        ("p_first", c_int), ("p_count", c_int),
        ("q_first", c_int), ("q_count", c_int),
        ("rank_padded", c_int64), ("splits", c_int64),
        ("path", c_int), ("rank_passes", c_int), ("f16", c_int), ("reserved_", c_int),
    ]


class Partials(ctypes.Structure):
    """Mirror of tlb200_partials_t."""

    _fields_ = [("data", c_void_p), ("splits", c_int64), ("split_stride", c_int64), ("ld", c_int64), ("rows", c_int64),
                ("rank", c_int64)]


_I64P = POINTER(c_int64)
_VPP = POINTER(c_void_p)
_INTP = POINTER(c_int)

# name -> (restype, argtypes); must list every symbol of include/tlb200.h
SIGNATURES = {
    "tlb200_version": (c_int, []),
    "tlb200_build_arch": (c_char_p, []),
    "tlb200_status_string": (c_char_p, [c_int]),
    "tlb200_last_path": (c_char_p, []),
    "tlb200_launch_count": (c_int64, []),
    "tlb200_source_hash": (c_char_p, []),
    "tlb200_unfold": (c_int, [c_void_p, _I64P, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tlb200_fold": (c_int, [c_void_p, _I64P, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tlb200_khatri_rao": (c_int, [_VPP, _I64P, _I64P, _I64P, c_int, c_int64, c_void_p, c_void_p, c_int, c_void_p,
                                  c_int64, c_void_p]),
    "tlb200_mttkrp_workspace_bytes": (c_size_t, [_I64P, c_int, c_int, c_int64, c_int, c_int]),
    "tlb200_mttkrp": (c_int, [c_void_p, _I64P, c_int, c_int, _VPP, _I64P, _I64P, c_int64, c_void_p, c_int, c_void_p,
                              c_int64, c_void_p, c_size_t, c_int, c_void_p]),
    "tlb200_mttkrp_plan": (c_int, [_I64P, c_int, c_int, c_int64, c_int, c_int, POINTER(MttkrpPlan)]),
    "tlb200_mode_dot_workspace_bytes": (c_size_t, [_I64P, c_int, c_int, c_int64, c_int, c_int]),
    "tlb200_mttkrp_from_ttm_workspace_bytes": (c_size_t, [_I64P, c_int, c_int, c_int64, c_int]),
    "tlb200_mttkrp_from_ttm": (c_int, [c_void_p, _I64P, c_int, c_int, _VPP, _I64P, _I64P, c_int64, c_void_p, c_int, c_void_p,
                                       c_int64, c_void_p, c_size_t, c_void_p]),
    "tlb200_mode_dot": (c_int, [c_void_p, _I64P, c_int, c_int, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p,
                                c_void_p, c_size_t, c_int, c_void_p]),
    "tlb200_multi_mode_dot_workspace_bytes": (c_size_t, [_I64P, c_int, _INTP, _I64P, c_int, c_int, c_int]),
    "tlb200_multi_mode_dot": (c_int, [c_void_p, _I64P, c_int, _INTP, _VPP, _I64P, _I64P, _I64P, c_int, c_int, c_void_p,
                                      c_void_p, c_size_t, c_int, c_void_p]),
    "tlb200_gram_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int]),
    "tlb200_gram": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "tlb200_cp_update": (c_int, [_VPP, c_int, c_int, c_int64, c_void_p, c_double, c_void_p, c_int64, c_int64, c_int,
                                 c_void_p, c_int64, c_void_p]),
    "tlb200_cp_update_gram_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int]),
    "tlb200_cp_update_gram": (c_int, [_VPP, c_int, c_int, c_int64, c_void_p, c_double, c_void_p, c_int64, c_int64, c_int,
                                      c_void_p, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "tlb200_cp_error": (c_int, [_VPP, c_int, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64,
                                c_void_p, c_int, c_void_p, c_void_p]),
    "tlb200_sumsq_workspace_bytes": (c_size_t, [c_int64, c_int]),
    "tlb200_sumsq": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "tlb200_nncp_update": (c_int, [_VPP, c_int, c_int, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64,
                                   c_double, c_int, c_void_p]),
    "tlb200_cp_to_tensor": (c_int, [_VPP, _I64P, _I64P, _I64P, c_int, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "tlb200_cp_impute_workspace_bytes": (c_size_t, [_I64P, c_int]),
    "tlb200_cp_impute": (c_int, [c_void_p, c_void_p, _VPP, _I64P, _I64P, _I64P, c_int, c_int64, c_void_p, c_int, c_void_p,
                                 c_void_p, c_void_p, c_size_t, c_void_p]),
    "tlb200_orthonormalize_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "tlb200_orthonormalize": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int, c_void_p, c_int64, c_int,
                                      c_void_p, c_size_t, c_void_p]),
    "tlb200_symeig": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_int64, c_void_p]),
    "tlb200_hals_workspace_bytes": (c_size_t, [c_int64]),
    "tlb200_hals_update": (c_int, [_VPP, c_int, c_int, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int64,
                                   c_int64, c_int64, c_int, c_double, POINTER(c_double), POINTER(c_double), c_double, c_int,
                                   c_void_p, c_void_p, c_size_t, c_void_p]),
    "tlb200_subspace_iterate_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "tlb200_subspace_iterate": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int, c_void_p, c_size_t,
                                        c_void_p]),
    "tlb200_comm_buffer_bytes": (c_size_t, [c_int, c_size_t]),
    "tlb200_comm_alloc": (c_int, [c_size_t, POINTER(c_void_p), c_void_p]),
    "tlb200_comm_open": (c_int, [c_void_p, POINTER(c_void_p)]),
    "tlb200_comm_close": (c_int, [c_void_p]),
    "tlb200_comm_free": (c_int, [c_void_p]),
    "tlb200_allreduce_oneshot": (c_int, [c_void_p, c_void_p, c_int64, c_int, _VPP, c_int, c_int, c_size_t, c_void_p]),
    "tlb200_mttkrp_partials": (c_int, [c_void_p, _I64P, c_int, c_int, _VPP, _I64P, _I64P, c_int64, c_void_p, c_int, c_void_p,
                                       c_size_t, c_int, POINTER(Partials), c_void_p]),
    "tlb200_mttkrp_from_ttm_partials": (c_int, [c_void_p, _I64P, c_int, c_int, _VPP, _I64P, _I64P, c_int64, c_void_p, c_int,
                                                c_void_p, c_size_t, POINTER(Partials), c_void_p]),
    "tlb200_tensor_absmax": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "tlb200_hint_tensor_absmax": (c_int, [c_void_p, c_void_p]),
    "tlb200_cp_update_fused": (c_int, [_VPP, c_int, c_int, c_int64, c_void_p, c_double, POINTER(Partials), c_int, c_void_p,
                                       c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                       c_void_p]),
    "tlb200_cp_error_iprod": (c_int, [_VPP, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "tlb200_allreduce_partials": (c_int, [POINTER(Partials), c_void_p, c_int64, c_void_p, c_int, _VPP, c_int, c_int, c_size_t,
                                          c_void_p]),
}

_lib = None


def source_hash() -> str:
    """sha256 over the CUDA sources and headers the library is built from (file names + contents, sorted).
    build.py bakes it into the .so (tlb200_source_hash); load() refuses a library built from other sources."""
    import hashlib
    csrc = os.path.join(_HERE, "csrc")
    files = sorted(os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh", ".h")))
    files.append(os.path.join(os.path.dirname(_HERE), "include", "tlb200.h"))
    h = hashlib.sha256()
    for path in files:
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:32]


def load() -> ctypes.CDLL:
    """Load libtlb200.so (once). Raises RuntimeError if it is missing — never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"tensorly_b200: native library not found at {LIB_PATH}. Build it with "
            "`python -m tensorly_b200.build` (needs nvcc); there is no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    built_from = lib.tlb200_source_hash().decode()
    if os.environ.get("TLB200_SKIP_HASH_CHECK", "0") != "1" and built_from != source_hash():
        raise RuntimeError(
            f"tensorly_b200: {LIB_PATH} was built from other sources (stamp {built_from}, tree {source_hash()}); "
            "rebuild it with `python -m tensorly_b200.build`.")
    _lib = lib
    return lib


def i64_array(values):
    values = list(values)
    return (c_int64 * max(1, len(values)))(*values)


def int_array(values):
    values = list(values)
    return (c_int * max(1, len(values)))(*values)


def ptr_array(ptrs):
    ptrs = list(ptrs)
    return (c_void_p * max(1, len(ptrs)))(*[c_void_p(p) if p else c_void_p(None) for p in ptrs])


def check(status: int, what: str) -> None:
    """Map a tlb200 status to the exception the reference raises for the same misuse."""
    if status == TLB200_OK:
        return
    msg = f"{what}: {load().tlb200_status_string(status).decode()} (status {status})"
    if status == TLB200_EINVAL:
        raise ValueError(msg)
    raise RuntimeError(msg)

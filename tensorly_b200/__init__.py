"""tensorly_b200 — a B200-native (sm_100a) backend for TensorLy's dense-decomposition hot path.

Public surface (same names and signatures as the reference functions they replace):

    unfold, fold                                   tensorly/base.py
    khatri_rao, unfolding_dot_khatri_rao           tensorly/tenalg/core_tenalg
    mode_dot, multi_mode_dot                       tensorly/tenalg/core_tenalg
    parafac, non_negative_parafac                  tensorly/decomposition (own sweep loop,
                                                   CUDA-graphed, optionally mode-sharded)
    register(), use()                              plug the "b200" tenalg backend into an
                                                   unmodified TensorLy

All arithmetic runs in hand-written CUDA kernels behind the C ABI of include/tlb200.h;
there is no CPU fallback — a missing libtlb200.so raises at first use.
"""
from ._ops import (cp_error, cp_impute, cp_to_tensor, cp_update, fold, orthonormalize, get_kernel_path, gram, khatri_rao, last_kernel_path, launch_count, mode_dot,
                   hals_nnls, hals_update, mttkrp_from_ttm, mttkrp_plan, multi_mode_dot, nncp_update, release_workspaces, set_kernel_path, subspace_iterate, sumsq, symeig, unfold,
                   unfolding_dot_khatri_rao, tensor_absmax, RangeHint)
from .backend import BACKEND_NAME, import_tensorly, register, set_dimension_tree, use
from .cp_als import CPALS, CPResult, non_negative_parafac, non_negative_parafac_hals, parafac, shard_bounds
from .solve import fast_solve, use_default_solve, use_fast_solve
from .svd import gram_svd, use_default_svd, use_gram_svd
from .tucker_hooi import HOOI, partial_tucker, tucker, tucker_to_tensor

__version__ = "0.1.0"
__all__ = [
    "tensor_absmax", "RangeHint",
    "unfold", "fold", "khatri_rao", "unfolding_dot_khatri_rao", "mode_dot", "multi_mode_dot",
    "parafac", "non_negative_parafac", "non_negative_parafac_hals", "hals_nnls", "hals_update", "CPALS", "CPResult", "shard_bounds",
    "tucker", "partial_tucker", "tucker_to_tensor", "HOOI", "orthonormalize", "symeig", "subspace_iterate",
    "gram", "cp_update", "nncp_update", "cp_error", "sumsq", "mttkrp_plan", "mttkrp_from_ttm", "cp_to_tensor", "cp_impute",
    "set_kernel_path", "get_kernel_path", "last_kernel_path", "launch_count", "release_workspaces",
    "register", "use", "set_dimension_tree", "import_tensorly", "BACKEND_NAME",
    "gram_svd", "use_gram_svd", "use_default_svd", "fast_solve", "use_fast_solve", "use_default_solve",
]

/*
 * tlb200.h — C ABI of the B200-native TensorLy tenalg hot path.
 *
 * Every entry point takes plain device pointers, shapes and a CUDA stream
 * (passed as void* so that this header needs no CUDA include).  The caller
 * owns every buffer (inputs, output, workspace); the callee allocates nothing,
 * never synchronises the device and launches only on `stream`.
 *
 * Return value: 0 on success, a negative tlb200_status otherwise.  The Python
 * shim (tensorly_b200/_ops.py) maps TLB200_EINVAL to ValueError — the same
 * exception the reference raises for shape mismatches — and everything else to
 * RuntimeError.
 *
 * Each function cites the reference interface it replaces (paths relative to
 * the tensorly/tensorly checkout, v0.9.0).
 */
#ifndef TLB200_H_
#define TLB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TLB200_MAX_NDIM 8

typedef enum {
    TLB200_OK = 0,
    TLB200_EINVAL = -1,    /* bad shape / mode / rank / dtype            */
    TLB200_EWORKSPACE = -2,/* workspace too small                         */
    TLB200_ECUDA = -3,     /* a CUDA runtime call or launch failed        */
    TLB200_EUNSUPPORTED = -4
} tlb200_status;

typedef enum { TLB200_F32 = 0, TLB200_F64 = 1 } tlb200_dtype;

/* Which kernel family a call may use.  AUTO picks the tcgen05 (tensor-core,
 * 3xTF32 error-compensated) path when dtype/shape/alignment allow it and the
 * SIMT path otherwise; both are hand-written sm_100a kernels, there is no
 * host fallback. */
typedef enum { TLB200_PATH_AUTO = 0, TLB200_PATH_SIMT = 1, TLB200_PATH_TCGEN05 = 2 } tlb200_path;

/* Library / build introspection. */
int         tlb200_version(void);
const char* tlb200_build_arch(void);            /* "sm_100a" */
const char* tlb200_status_string(int status);
/* Name of the kernel family the last call on this host thread dispatched to
 * ("simt", "tcgen05", "copy", ...).  Used by tests to prove which path ran. */
const char* tlb200_last_path(void);
/* Number of kernel launches this library has issued so far in this process (host-side
 * count; graph replays re-launch the captured kernels without passing through here). */
int64_t     tlb200_launch_count(void);
/* Stamp of the sources this binary was compiled from (sha256 prefix over csrc + this header, written by
 * tensorly_b200/build.py); the Python loader refuses a library whose stamp differs from the tree beside it. */
const char* tlb200_source_hash(void);

/* ---------------------------------------------------------------------------
 * unfold — replaces tensorly.base.unfold (tensorly/base.py:39-53):
 *   out[i_mode, (i_0..i_{mode-1}, i_{mode+1}..i_{N-1})] = x[i_0..i_{N-1}]
 * x: C-contiguous N-way tensor; out: C-contiguous (shape[mode], prod(others)).
 * Pure data movement, bit-exact.
 * ------------------------------------------------------------------------- */
int tlb200_unfold(const void* x, const int64_t* shape, int ndim, int mode,
                  int dtype, void* out, void* stream);

/* fold — replaces tensorly.base.fold (tensorly/base.py:56-79); inverse of unfold.
 * unfolded: C-contiguous (shape[mode], prod(others)); out: C-contiguous `shape`. */
int tlb200_fold(const void* unfolded, const int64_t* shape, int ndim, int mode,
                int dtype, void* out, void* stream);

/* ---------------------------------------------------------------------------
 * khatri_rao — replaces tensorly.tenalg.core_tenalg.khatri_rao
 * (tensorly/tenalg/core_tenalg/_khatri_rao.py:9-109).
 *   out[(i_0,..,i_{m-1}), r] = (((w[r]*M_0[i_0,r]) * M_1[i_1,r]) * ...) * mask[row]
 * evaluated as a left fold with one IEEE rounding per multiply (no FMA, no
 * reassociation), so the result is bit-identical to the reference's chain of
 * broadcast multiplies.  First matrix slowest, last fastest.
 * mats[i] has rows[i] rows and `rank` columns, element (r,c) at
 * mats[i][r*row_stride[i] + c*col_stride[i]] (strides in elements).
 * weights (rank,) and mask (prod(rows),) may be NULL.  out: (prod(rows), rank)
 * with row stride out_ld >= rank.
 * ------------------------------------------------------------------------- */
int tlb200_khatri_rao(const void* const* mats, const int64_t* rows,
                      const int64_t* row_stride, const int64_t* col_stride,
                      int nmats, int64_t rank, const void* weights,
                      const void* mask, int dtype, void* out, int64_t out_ld,
                      void* stream);

/* ---------------------------------------------------------------------------
 * MTTKRP — replaces tensorly.tenalg.core_tenalg.unfolding_dot_khatri_rao
 * (tensorly/tenalg/core_tenalg/mttkrp.py:9-50):
 *   out[i_mode, r] = sum_{i_n, n != mode} x[i_0..i_{N-1}] * w[r] * prod_{n != mode} F_n[i_n, r]
 * x: C-contiguous N-way tensor, streamed from HBM exactly once; the Khatri-Rao
 * matrix and the unfolding are never materialised.  factors[n] is (shape[n], rank)
 * with element strides; factors[mode] is ignored (may be NULL).  weights may be
 * NULL.  out: (shape[mode], rank) with row stride out_ld.
 * `workspace` must hold tlb200_mttkrp_workspace_bytes(...) bytes (256-byte aligned).
 * ------------------------------------------------------------------------- */
size_t tlb200_mttkrp_workspace_bytes(const int64_t* shape, int ndim, int mode,
                                     int64_t rank, int dtype, int path);

int tlb200_mttkrp(const void* x, const int64_t* shape, int ndim, int mode,
                  const void* const* factors, const int64_t* f_row_stride,
                  const int64_t* f_col_stride, int64_t rank, const void* weights,
                  int dtype, void* out, int64_t out_ld, void* workspace,
                  size_t workspace_bytes, int path, void* stream);

/* The (A, J, B) streaming plan tlb200_mttkrp uses for a shape/mode: the tensor is
 * viewed as X[a, j, b] with strides (sa, sj, sb) and the KR rows factor as
 * P[a, :] * Q[b, :].  Host-only (no device work); exposed so the planning logic is
 * testable without a GPU. */
typedef struct {
    int64_t A, J, B;          /* extents                                    */
    int64_t sa, sj, sb;       /* element strides of the view                */
    int     p_first, p_count; /* modes folded into the P table              */
    int     q_first, q_count; /* modes folded into the Q table              */
    int64_t rank_padded;      /* column count of the P/Q/partials tables    */
    int64_t splits;           /* split-K factor (deterministic 2-pass sum)  */
    int     path;             /* resolved tlb200_path                       */
    int     rank_passes;      /* passes over the tensor: 1, or ceil(rank/64) column blocks
                                 on the tcgen05 path when rank > 64 (the other fields then
                                 describe the first pass)                   */
    int     f16;              /* 1: planned for the fp16-split engine (a range hint is registered
                                 for the tensor, see tlb200_hint_tensor_absmax); the shape-only
                                 tlb200_mttkrp_plan always reports the 3xTF32 plan (0) */
    int     reserved_;
} tlb200_mttkrp_plan_t;

int tlb200_mttkrp_plan(const int64_t* shape, int ndim, int mode, int64_t rank,
                       int dtype, int path, tlb200_mttkrp_plan_t* plan);

/* Dimension-tree reuse inside one ALS sweep (the caller side of the MTTKRP path,
 * tensorly/decomposition/_cp.py:407-428 calls unfolding_dot_khatri_rao once per mode):
 * with T = mode_dot(x, F_{N-1}^T, N-1), shape (I_0, .., I_{N-2}, rank), C-contiguous,
 *   tlb200_mttkrp_from_ttm(t, mode)[j, r] = w_r * sum T[.., j, .., r] * prod_{n != mode, n < N-1} F_n[i_n, r]
 * equals tlb200_mttkrp(x, mode) for every mode < N-1 as long as F_{N-1} is unchanged, at
 * rank / I_{N-1} of the memory traffic.  lead_shape = (I_0, .., I_{N-2}), nlead = N-1 >= 2;
 * factors / strides have nlead entries (entry `mode` ignored). */
size_t tlb200_mttkrp_from_ttm_workspace_bytes(const int64_t* lead_shape, int nlead, int mode,
                                              int64_t rank, int dtype);

int tlb200_mttkrp_from_ttm(const void* t, const int64_t* lead_shape, int nlead, int mode,
                           const void* const* factors, const int64_t* f_row_stride,
                           const int64_t* f_col_stride, int64_t rank, const void* weights,
                           int dtype, void* out, int64_t out_ld, void* workspace,
                           size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------
 * mode_dot — replaces tensorly.tenalg.core_tenalg.mode_dot for a matrix operand
 * (tensorly/tenalg/core_tenalg/n_mode_product.py:5-76):
 *   out[.., i, ..] = sum_j m[i, j] * x[.., j, ..]      (contracting `mode`)
 * m is (rows_out, shape[mode]) addressed as m[i*m_row_stride + j*m_col_stride], so
 * `transpose=True` in the reference is a stride swap on the caller's side.  A vector
 * operand is rows_out == 1 (the caller drops the mode).  x and out are C-contiguous;
 * out has shape[mode] replaced by rows_out.
 * ------------------------------------------------------------------------- */
size_t tlb200_mode_dot_workspace_bytes(const int64_t* shape, int ndim, int mode,
                                       int64_t rows_out, int dtype, int path);

int tlb200_mode_dot(const void* x, const int64_t* shape, int ndim, int mode,
                    const void* m, int64_t rows_out, int64_t m_row_stride,
                    int64_t m_col_stride, int dtype, void* out, void* workspace,
                    size_t workspace_bytes, int path, void* stream);

/* ---------------------------------------------------------------------------
 * multi_mode_dot — replaces tensorly.tenalg.core_tenalg.multi_mode_dot for matrix
 * operands (tensorly/tenalg/core_tenalg/n_mode_product.py:79-135): the chain
 *   x  x_{modes[0]} m_0  x_{modes[1]} m_1 ...
 * over `nmats` DISTINCT, ascending modes (the caller has already applied `skip`,
 * sorted by mode and resolved `transpose` into strides).  Intermediates live in
 * `workspace`.  out is C-contiguous with each contracted extent replaced by
 * rows_out[k].
 * ------------------------------------------------------------------------- */
size_t tlb200_multi_mode_dot_workspace_bytes(const int64_t* shape, int ndim,
                                             const int* modes, const int64_t* rows_out,
                                             int nmats, int dtype, int path);

int tlb200_multi_mode_dot(const void* x, const int64_t* shape, int ndim,
                          const int* modes, const void* const* mats,
                          const int64_t* rows_out, const int64_t* m_row_stride,
                          const int64_t* m_col_stride, int nmats, int dtype,
                          void* out, void* workspace, size_t workspace_bytes,
                          int path, void* stream);

/* ---------------------------------------------------------------------------
 * CP-ALS normal equations — replaces the caller-side lines of
 * tensorly.decomposition._cp.parafac (tensorly/decomposition/_cp.py:411-428):
 *
 * tlb200_gram: G = F^T F for F (rows, rank) [rank x rank, row-major, ld = rank].
 *   Deterministic: per-block partial sums are written to workspace and summed in a
 *   fixed order by the last block.
 *
 * tlb200_cp_update: V = (w w^T) o prod_{i != mode} G_i + l2_reg * I ;
 *                   F_mode = solve(V^T, M^T)^T   (LU with partial pivoting, as
 *                   numpy/torch `solve` does), for `rows` rows of the MTTKRP M.
 *   grams: nmodes pointers to rank x rank Gram matrices (entry `mode` ignored).
 *   out may alias m.
 * ------------------------------------------------------------------------- */
size_t tlb200_gram_workspace_bytes(int64_t rows, int64_t rank, int dtype);

int tlb200_gram(const void* f, int64_t rows, int64_t rank, int64_t row_stride,
                int64_t col_stride, int dtype, void* gram, void* workspace,
                size_t workspace_bytes, void* stream);

int tlb200_cp_update(const void* const* grams, int nmodes, int mode, int64_t rank,
                     const void* weights, double l2_reg, const void* m, int64_t m_ld,
                     int64_t rows, int dtype, void* out, int64_t out_ld, void* stream);

/* tlb200_cp_update followed by tlb200_gram of the updated rows, in one launch: the solved rows
 * are still in shared memory when their Gram partials are formed (no second pass over the
 * factor, no extra launches).  gram_out (rank x rank) may be grams[mode].  The first 4 bytes of
 * `workspace` are a ticket counter: they must be zero before the first call and are left zero
 * by every call, so one zero-initialised workspace can be reused forever on one stream.
 * Replaces tensorly/decomposition/_cp.py:411-428 plus the later recomputation of
 * `tl.dot(tl.conj(tl.transpose(factor)), factor)` for this factor at _cp.py:413-416. */
size_t tlb200_cp_update_gram_workspace_bytes(int64_t rows, int64_t rank, int dtype);

int tlb200_cp_update_gram(const void* const* grams, int nmodes, int mode, int64_t rank,
                          const void* weights, double l2_reg, const void* m, int64_t m_ld,
                          int64_t rows, int dtype, void* out, int64_t out_ld, void* gram_out,
                          void* workspace, size_t workspace_bytes, void* stream);

/* Split-K partials of an MTTKRP, left unsummed for tlb200_cp_update_fused: element (split s, row i, column r) at
 * data[s * split_stride + i * ld + r]; the MTTKRP is their sum over s in index order.  `data` points into the
 * workspace of the call that produced it and stays valid until that workspace is reused. */
typedef struct {
    const void* data;
    int64_t splits, split_stride, ld;
    int64_t rows, rank;
} tlb200_partials_t;

/* tlb200_mttkrp / tlb200_mttkrp_from_ttm without their final reduction launch: same arguments minus `out`, the
 * partials are described in *partials.  TLB200_EUNSUPPORTED for rank > 64 on the tensor-core path (several passes). */
int tlb200_mttkrp_partials(const void* x, const int64_t* shape, int ndim, int mode,
                           const void* const* factors, const int64_t* f_row_stride,
                           const int64_t* f_col_stride, int64_t rank, const void* weights,
                           int dtype, void* workspace, size_t workspace_bytes, int path,
                           tlb200_partials_t* partials, void* stream);

int tlb200_mttkrp_from_ttm_partials(const void* t, const int64_t* lead_shape, int nlead, int mode,
                                    const void* const* factors, const int64_t* f_row_stride,
                                    const int64_t* f_col_stride, int64_t rank, const void* weights,
                                    int dtype, void* workspace, size_t workspace_bytes,
                                    tlb200_partials_t* partials, void* stream);

/* tlb200_cp_update_gram whose right-hand sides are those partials: they are summed (split order) while the LU runs,
 * so the reduction costs neither a launch nor time on the critical path.  m_out (optional, rows x rank) receives the
 * summed MTTKRP; iprod_out (optional device scalar) receives <M, F_new> = sum(M o F_new), the inner-product term of
 * the fast error (tensorly/decomposition/_cp.py:222) — tlb200_cp_error_iprod finishes the error from it, or, when
 * `mode` is the LAST mode, this launch does it itself: with err_out (3 scalars, needs iprod_out and norm_x2) the
 * tail also forms ||cp||^2 from the Gram matrices and writes [rel_error, <M, F>, ||cp||^2] like tlb200_cp_error.
 * Workspace: tlb200_cp_update_gram_workspace_bytes, same zero-ticket convention. */
int tlb200_cp_update_fused(const void* const* grams, int nmodes, int mode, int64_t rank,
                           const void* weights, double l2_reg, const tlb200_partials_t* m, int dtype,
                           void* out, int64_t out_ld, void* gram_out, void* m_out, int64_t m_out_ld,
                           void* iprod_out, const void* norm_x2, void* err_out, void* workspace,
                           size_t workspace_bytes, void* stream);

int tlb200_cp_error_iprod(const void* const* grams, int nmodes, int64_t rank, const void* weights,
                          const void* iprod, const void* norm_x2, int dtype, void* err_out,
                          void* stream);

/* Fast CP reconstruction error — replaces error_calc's MTTKRP shortcut
 * (tensorly/decomposition/_cp.py:217-225 with cp_norm, tensorly/cp_tensor.py:614-644):
 *   iprod = sum(M_last o F_last);  norm_cp^2 = sum_{r,s} w_r w_s prod_n G_n[r,s];
 *   err_out[0] = sqrt(|norm_x2 + norm_cp^2 - 2 iprod|) / sqrt(norm_x2)
 *   err_out[1] = iprod, err_out[2] = norm_cp^2.
 * norm_x2 points to a device scalar holding ||X||^2.  err_out: 3 scalars of `dtype`. */
int tlb200_cp_error(const void* const* grams, int nmodes, int64_t rank,
                    const void* weights, const void* m_last, int64_t m_ld,
                    const void* f_last, int64_t f_row_stride, int64_t f_col_stride,
                    int64_t rows, const void* norm_x2, int dtype, void* err_out,
                    void* stream);

/* Range hint for the fp16-split tensor-core engine.  The 3xTF32 engine needs only the data; its fp16 sibling (half
 * the tensor-core instructions and operand traffic per tensor byte — what rank 33..64 is bound by under the power
 * cap) computes on x * 2^k with max |x| mapped into fp16 range, so it needs max |x|.  tlb200_tensor_absmax writes it
 * to a device float (fp32 tensors only; one HBM pass, no host sync); tlb200_hint_tensor_absmax registers that device
 * scalar for the tensor whose base pointer is `x` (NULL withdraws the hint).  While a hint is registered,
 * tlb200_mttkrp / tlb200_mttkrp_partials / tlb200_mode_dot calls on that pointer take the fp16 engine (same
 * results to ~1e-6: every element keeps 22 significant bits relative to max |x|; elements below 2^-28 max |x| keep
 * absolute accuracy 2^-50 max |x|).  The caller promises the scalar stays valid and >= max |x| for as long as the
 * hint is registered (the ALS drivers: the tensor is constant over a decomposition).  No reference counterpart. */
int tlb200_tensor_absmax(const void* x, int64_t n, int dtype, void* absmax_out, void* stream);
int tlb200_hint_tensor_absmax(const void* x, const void* absmax_device);

/* Sum of squares of a contiguous array (||X||^2) into a device scalar of `dtype`
 * (accumulated in double).  Replaces tl.norm(tensor, 2)**2 at _cp.py:350. */
size_t tlb200_sumsq_workspace_bytes(int64_t n, int dtype);
int tlb200_sumsq(const void* x, int64_t n, int dtype, void* out, void* workspace,
                 size_t workspace_bytes, void* stream);

/* Non-negative CP multiplicative update — replaces
 * tensorly/decomposition/_nn_cp.py:131-136:
 *   F[i,r] <- F[i,r] * max(M[i,r], eps) / max((F V)[i,r], eps)
 * with V = (w w^T) o prod_{i != mode} G_i formed exactly as in tlb200_cp_update.
 * f is updated in place (row stride f_ld). */
int tlb200_nncp_update(const void* const* grams, int nmodes, int mode, int64_t rank,
                       const void* weights, const void* m, int64_t m_ld, void* f,
                       int64_t f_ld, int64_t rows, double eps, int dtype, void* stream);

/* ---------------------------------------------------------------------------
 * CP reconstruction — replaces tensorly.cp_tensor.cp_to_tensor (tensorly/cp_tensor.py:433-485):
 *   out[i_0..i_{N-1}] = sum_r w_r * prod_n F_n[i_n, r]      (times mask[i_0..i_{N-1}] when mask != NULL)
 * factors[n] is (shape[n], rank) with element strides; weights (rank,) and mask (C-contiguous, `shape`) may be
 * NULL.  out: C-contiguous `shape`, written exactly once; the Khatri-Rao matrix of the reference is never formed.
 * 2 <= ndim <= TLB200_MAX_NDIM (a vector is the caller's 2-way case with a 1 x rank factor of ones).
 * ------------------------------------------------------------------------- */
int tlb200_cp_to_tensor(const void* const* factors, const int64_t* shape,
                        const int64_t* f_row_stride, const int64_t* f_col_stride, int ndim,
                        int64_t rank, const void* weights, const void* mask, int dtype,
                        void* out, void* stream);

/* Masked-ALS imputation and error in ONE pass over the tensor — replaces the mask branch of error_calc
 * (tensorly/decomposition/_cp.py:195-207: cp_to_tensor, tensor*mask + rec*(1-mask), tl.norm twice):
 *   out   = x * mask + rec * (1 - mask)             (out may alias x; mask: 1 = observed, 0 = missing)
 *   stats = [ ||out - rec|| / ||out||,  ||out||^2,  ||out - rec||^2 ]      (3 device scalars of `dtype`,
 *           accumulated in double, fixed summation order)
 * rec is formed in registers and never stored. */
size_t tlb200_cp_impute_workspace_bytes(const int64_t* shape, int ndim);

int tlb200_cp_impute(const void* x, const void* mask, const void* const* factors,
                     const int64_t* shape, const int64_t* f_row_stride,
                     const int64_t* f_col_stride, int ndim, int64_t rank, const void* weights,
                     int dtype, void* out, void* stats, void* workspace, size_t workspace_bytes,
                     void* stream);

/* ---------------------------------------------------------------------------
 * Orthonormal basis of the span of a tall block — the truncated "SVD" step of the own HOOI driver, replacing
 * svd_interface(unfold(core_approximation, mode), n_eigenvecs=rank) inside the loop of
 * tensorly/decomposition/_tucker.py:197-201 (-> tensorly/tenalg/svd.py:211-235) together with warm-started
 * subspace iteration on the Gram matrix of the unfolding (tensorly_b200/tucker_hooi.py):
 *   out = Z R^{-1},  R^T R = Z^T Z   (Cholesky-QR; Gram, Cholesky factor and inverse in fp64; passes = 2 repeats
 *   it on the result, which makes `out` orthonormal to rounding whenever cond(Z)^2 < 1e16)
 * z: (rows, rank) with element strides, rank <= 64 <= ... <= rows; out: (rows, rank), row stride out_ld.
 * The first 8 bytes of `workspace` are a ticket counter and a status word: the counter must be zero before the
 * first call and is left zero; status (second int) is 1 when the block was numerically rank deficient.
 * ------------------------------------------------------------------------- */
size_t tlb200_orthonormalize_workspace_bytes(int64_t rows, int64_t rank);

int tlb200_orthonormalize(const void* z, int64_t rows, int64_t rank, int64_t row_stride,
                          int64_t col_stride, int dtype, void* out, int64_t out_ld, int passes,
                          void* workspace, size_t workspace_bytes, void* stream);

/* `steps` warm-started power steps  U <- orth(G U)  in place, fp64: G (n x n, symmetric, row stride g_ld), U (n x p,
 * p <= 64 <= n, row stride u_ld).  Two launches per step (GEMM block + Gram partial + Cholesky/inverse in the last
 * CTA; apply).  With tlb200_orthonormalize / tlb200_symeig this is the in-loop "SVD" of the own HOOI driver
 * (tensorly/decomposition/_tucker.py:197-201).  Workspace: first 8 bytes zero before the first call (left zero). */
size_t tlb200_subspace_iterate_workspace_bytes(int64_t n, int64_t p);

int tlb200_subspace_iterate(const void* g, int64_t n, int64_t g_ld, void* u, int64_t p, int64_t u_ld,
                            int steps, void* workspace, size_t workspace_bytes, void* stream);

/* Eigendecomposition of a small symmetric matrix (n <= 64): cyclic Jacobi with parallel ordering in fp64, one CTA.
 * evals (n) in DESCENDING order, evecs (n x n, row stride ldv) with the eigenvectors as columns.  The Rayleigh-Ritz
 * step of the HOOI subspace iteration (it replaces the ordering that the SVD of tensorly/tenalg/svd.py:211-235
 * provides); a is read as (a + a^T) / 2. */
int tlb200_symeig(const void* a, int64_t n, int64_t lda, int dtype, void* evals, void* evecs,
                  int64_t ldv, void* stream);

/* ---------------------------------------------------------------------------
 * HALS non-negative least squares in one launch — replaces tensorly.solvers.nnls.hals_nnls
 * (tensorly/solvers/nnls.py:5-175, rank loop :139-153) as called per mode by non_negative_parafac_hals
 * (tensorly/decomposition/_nn_cp.py:326-336):
 *   repeat <= n_iter_max times, for k = 0..rank-1:
 *     V[k,:] = max((UtM[k,:] - UtU[k,:] V + UtU[k,k] V[k,:] - sparsity) / (UtU[k,k] + 2 ridge), epsilon)
 *   until sum_k ||V - V_new[k,:]||^2 < tol * (its value in the first iteration)      (the reference's statistic)
 * UtU = (w w^T) o prod_{i != mode} grams[i] as in tlb200_cp_update, or grams[0] itself when mode < 0.
 * m = UtM^T and f = V^T (updated in place) are addressed as x[row * row_stride + k * col_stride], row < rows,
 * i.e. the MTTKRP and the factor of the CP driver are passed as they are.  sparsity / ridge: NULL = absent.
 * rank <= 64 (fp32) / 32 (fp64); rows <= a few 10^4 (one cooperative grid).  iters_out (device int, may be NULL)
 * receives the number of inner iterations run.
 * ------------------------------------------------------------------------- */
size_t tlb200_hals_workspace_bytes(int64_t rows);

int tlb200_hals_update(const void* const* grams, int nmodes, int mode, int64_t rank,
                       const void* weights, const void* m, int64_t m_row_stride,
                       int64_t m_col_stride, void* f, int64_t f_row_stride, int64_t f_col_stride,
                       int64_t rows, int n_iter_max, double tol, const double* sparsity,
                       const double* ridge, double epsilon, int dtype, void* iters_out,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------
 * One-shot all-reduce over NVLink peer memory (one process per GPU, one node) for the small partials of the
 * sharded CP-ALS sweep — the exchange step SURVEY.md 8(e) adds to tensorly/decomposition/_cp.py:407-428 (the
 * reference has no distributed path).  Every rank allocates one symmetric buffer (tlb200_comm_alloc), exports its
 * 64-byte CUDA IPC handle, maps the peers' buffers (tlb200_comm_open) and then calls tlb200_allreduce_oneshot
 * with the same sequence of sizes on every rank: one launch pushes the vector to every peer, publishes a flag,
 * waits for all flags and sums the `world` contributions in rank order (same bits on every rank; bounded waits
 * trap instead of hanging).  in may equal out.  count * sizeof(dtype) <= max_payload_bytes.
 * ------------------------------------------------------------------------- */
size_t tlb200_comm_buffer_bytes(int world, size_t max_payload_bytes);
int tlb200_comm_alloc(size_t bytes, void** ptr, void* ipc_handle_out);
int tlb200_comm_open(const void* ipc_handle, void** peer_ptr);
int tlb200_comm_close(void* peer_ptr);
int tlb200_comm_free(void* ptr);
int tlb200_allreduce_oneshot(const void* in, void* out, int64_t count, int dtype, void* const* bufs,
                             int world, int rank, size_t max_payload_bytes, void* stream);
/* The same exchange with the split-K reduction of the MTTKRP fused into its push phase (compute + collective in
 * one kernel): the local contribution is the ordered sum over splits of *m, followed by tail_count plain elements
 * of `tail` (may be NULL; the R x R Gram partial that shares the exchange).  out: contiguous. */
int tlb200_allreduce_partials(const tlb200_partials_t* m, const void* tail, int64_t tail_count, void* out,
                              int dtype, void* const* bufs, int world, int rank,
                              size_t max_payload_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TLB200_H_ */

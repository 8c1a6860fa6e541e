"""CPU stand-in for tensorly_b200.cp_als.CudaOps built on the oracle — TEST ONLY.
Lets the sharded CP-ALS driver's host logic (slab partition, which partials are
all-reduced, error assembly) run under gloo on a box without a GPU."""
import numpy as np
import torch

from oracle import oracle as O


def _np(t):
    return None if t is None else t.detach().cpu().numpy()


def _form_v(grams, mode, weights, l2):
    rank = grams[0].shape[0]
    v = np.ones((rank, rank), dtype=_np(grams[0]).dtype)
    for i, g in enumerate(grams):
        if i != mode:
            v = v * _np(g)
    if l2:
        v = v + np.eye(rank, dtype=v.dtype) * l2
    w = _np(weights)
    return w[:, None] * v * w[None, :]


class OracleOps:
    supports_graphs = False

    @staticmethod
    def mttkrp(x, cp, mode, out=None):
        w, fs = cp
        m = torch.from_numpy(np.ascontiguousarray(O.unfolding_dot_khatri_rao(_np(x), (_np(w), [_np(f) for f in fs]), mode)))
        return m if out is None else out.copy_(m)

    @staticmethod
    def mode_dot(x, m, mode, transpose=False):
        return torch.from_numpy(np.ascontiguousarray(O.mode_dot(_np(x), _np(m), mode, transpose=transpose)))

    @staticmethod
    def mttkrp_from_ttm(t, cp, mode, out=None):
        """MTTKRP of a mode before the last from T = X x_last F_last^T: the same sum, written as an einsum
        over T (leading modes i_0..i_{N-2}, trailing r) and the other leading factors."""
        w, fs = cp
        t = _np(t)
        nlead = t.ndim - 1
        letters = "abcdefg"[:nlead]
        operands, subs = [t], [letters + "r"]
        for i in range(nlead):
            if i != mode:
                operands.append(_np(fs[i]))
                subs.append(letters[i] + "r")
        res = np.einsum(",".join(subs) + "->" + letters[mode] + "r", *operands)
        if w is not None:
            res = res * _np(w)[None, :]
        res = torch.from_numpy(np.ascontiguousarray(res.astype(t.dtype)))
        return res if out is None else out.copy_(res)

    @staticmethod
    def gram(f, out=None):
        g = torch.from_numpy(_np(f).T @ _np(f))
        return g if out is None else out.copy_(g)

    @staticmethod
    def cp_update(grams, mode, weights, m, l2_reg=0.0, out=None):
        v = _form_v(grams, mode, weights, l2_reg)
        f = torch.from_numpy(np.ascontiguousarray(np.linalg.solve(v.T, _np(m).T).T))
        return f if out is None else out.copy_(f)

    @staticmethod
    def nncp_update(grams, mode, weights, m, factor, eps):
        v = _form_v(grams, mode, weights, 0.0)
        f = _np(factor)
        factor.copy_(torch.from_numpy(f * np.clip(_np(m), eps, None) / np.clip(f @ v, eps, None)))
        return factor

    @staticmethod
    def hals_update(grams, mode, weights, m, factor, n_iter_max=100, tol=1e-8, sparsity_coefficient=None,
                    ridge_coefficient=None, epsilon=0.0, iters_out=None):
        v = _form_v(grams, mode, weights, 0.0)
        new = O.hals_nnls(_np(m).T, v, _np(factor).T, n_iter_max=n_iter_max, tol=tol,
                          sparsity_coefficient=sparsity_coefficient, ridge_coefficient=ridge_coefficient, epsilon=epsilon)
        factor.copy_(torch.from_numpy(np.ascontiguousarray(new.T)))
        return factor

    @staticmethod
    def cp_error(grams, weights, m_last, f_last, norm_x2, out=None):
        w = _np(weights)
        ncp = np.ones_like(_np(grams[0]))
        for g in grams:
            ncp = ncp * _np(g)
        ncp = float(np.sum(ncp * np.outer(w, w)))
        iprod = float(np.sum(_np(m_last) * _np(f_last)))
        nx2 = float(norm_x2[0])
        vals = torch.tensor([np.sqrt(abs(nx2 + ncp - 2 * iprod)) / np.sqrt(nx2), iprod, ncp], dtype=m_last.dtype)
        return vals if out is None else out.copy_(vals)

    @staticmethod
    def cp_impute(x, mask, cp, out=None, stats=None):
        w, fs = cp
        new, nrm, unnorm = O.cp_impute(_np(x), _np(mask), (_np(w), [_np(f) for f in fs]))
        new = torch.from_numpy(np.ascontiguousarray(new.astype(_np(x).dtype)))
        vals = torch.tensor([unnorm / nrm, nrm ** 2, unnorm ** 2], dtype=x.dtype)
        out = new if out is None else out.copy_(new)
        return out, (vals if stats is None else stats.copy_(vals))

    @staticmethod
    def sumsq(x, out=None):
        v = torch.tensor([float(np.sum(_np(x).astype(np.float64) ** 2))], dtype=x.dtype)
        return v if out is None else out.copy_(v)


    # ---- HOOI pieces (tensorly_b200.tucker_hooi.CudaOps stand-ins) ----
    @staticmethod
    def multi_mode_dot(x, mats, modes=None, skip=None, transpose=False):
        return torch.from_numpy(np.ascontiguousarray(O.multi_mode_dot(_np(x), [_np(m) for m in mats], modes=modes, skip=skip,
                                                                      transpose=transpose)))

    @staticmethod
    def unfold(x, mode, contiguous=False):
        return torch.from_numpy(np.ascontiguousarray(O.unfold(_np(x), mode)))

    @staticmethod
    def orthonormalize(z, out=None, passes=2):
        """Cholesky-QR in fp64, like tlb200_orthonormalize."""
        zz = _np(z).astype(np.float64)
        r = np.linalg.cholesky(zz.T @ zz).T
        q = torch.from_numpy(np.ascontiguousarray(zz @ np.linalg.inv(r)).astype(_np(z).dtype))
        return q if out is None else out.copy_(q)

    @staticmethod
    def symeig(a):
        w, v = np.linalg.eigh(0.5 * (_np(a) + _np(a).T))
        return torch.from_numpy(w[::-1].copy()), torch.from_numpy(np.ascontiguousarray(v[:, ::-1]))

    @staticmethod
    def subspace_iterate(g, u, steps):
        gg, uu = _np(g).astype(np.float64), _np(u).astype(np.float64)
        for _ in range(steps):
            z = gg @ uu
            r = np.linalg.cholesky(z.T @ z).T
            uu = z @ np.linalg.inv(r)
        u.copy_(torch.from_numpy(np.ascontiguousarray(uu)))
        return u

    @staticmethod
    def eigh_top(g, p):
        w, v = np.linalg.eigh(_np(g).astype(np.float64))
        return torch.from_numpy(np.ascontiguousarray(v[:, ::-1][:, :int(p)]))

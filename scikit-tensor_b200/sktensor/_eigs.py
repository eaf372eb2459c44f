"""Leading eigenvectors of ``X_(n) X_(n)^T`` for a sparse tensor, on the device (reference: core.py:280-283,
``eigsh(Xn.dot(Xn.T), rank, which='LM')`` on the SciPy CSR unfolding).

The unfolding is never formed as a matrix product.  The non-zeros are viewed as a 2-way COO tensor
(row = i_n, column = id of the distinct combination of the other subscripts, numbered on the device by
``skt_compact_keys``) and the operator ``V -> X_(n) (X_(n)^T V)`` is two calls of the sparse MTTKRP kernel with
``V`` / ``W`` as the single factor (all columns of the block in one pass, segmented reduction, no atomics).

* modes up to 2048 long: the Gram matrix is assembled by applying the operator to identity blocks and handed to the
  cluster eigensolver (``skt_symeig_top``) -- exact, no iteration;
* longer modes: thick-restarted block Lanczos with full re-orthogonalisation: orthonormalisation and Rayleigh-Ritz
  through the tall Gram (FP64 tensor cores), the small eigenproblems through ``skt_symeig_top``, every tall update a GEMM
  (``skt_ttm_dense`` on the 2-way view).  Only the small coefficient matrices visit the host.
"""
import os

import numpy as np

from . import _device as dev
from . import _kernels as K

_BLOCK = 64
_DENSE_MAX = K.SYMEIG_MAX_N


class SparseGramOperator(object):
    """``apply(V) = X_(n) X_(n)^T V`` for device blocks ``V`` of at most 128 columns."""

    def __init__(self, X, n):
        t = dev.torch()
        self.n = int(n)
        self.rows = int(X.shape[n])
        others = [m for m in range(X.ndim) if m != n]
        nnz = len(X.vals)
        if nnz == 0 or not others:
            self.coo = None
            self.ncols = 1
            if nnz:                                         # a 1-way "tensor": X_(n) is a single column
                sub = t.from_numpy(np.ascontiguousarray(X.subs[n], dtype=np.int64)).to(dev.device())
                vals = dev.to_device(np.asarray(X.vals, dtype=np.float64))
                self.coo = K.CooHandle.from_device([sub, t.zeros_like(sub)], vals, (self.rows, 1))
            return
        # linear key of the other subscripts (any bijection gives the same Gram), numbered 0..ncols-1 on the device
        bits = 0
        key = None
        for m in others:
            sm = t.from_numpy(np.ascontiguousarray(X.subs[m], dtype=np.int64)).to(dev.device())
            key = sm if key is None else key * int(X.shape[m]) + sm
            bits += max(1, int(np.ceil(np.log2(max(int(X.shape[m]), 2)))))
        if bits > 63:
            raise ValueError('the column index of the unfolding does not fit 63 bits')
        ids, self.ncols = K.compact_keys(key, min(bits + 1, 64))
        sub = t.from_numpy(np.ascontiguousarray(X.subs[n], dtype=np.int64)).to(dev.device())
        vals = dev.to_device(np.asarray(X.vals, dtype=np.float64))
        self.coo = K.CooHandle.from_device([sub, ids.to(t.int64)], vals, (self.rows, self.ncols))

    def apply(self, V):
        b = int(V.shape[1])
        if self.coo is None:
            return dev.zeros((self.rows, b))
        W = K.mttkrp_coo(self.coo, [V, None], b, 1)          # X_(n)^T V   (ncols x b)
        return K.mttkrp_coo(self.coo, [None, W], b, 0)       # X_(n) W     (rows x b)

    def close(self):
        if self.coo is not None:
            self.coo.close()
            self.coo = None


class _DeviceOps(object):
    """The tall-matrix primitives of the block iteration on the device (tests drive the same loop with NumPy)."""

    def __init__(self, op):
        self.op = op
        self.t = dev.torch()

    def random(self, n, m, seed):
        return dev.to_device(np.random.RandomState(seed).rand(n, m) - 0.5)

    def apply(self, V):
        if V.shape[1] <= 128:
            return self.op.apply(V.contiguous())
        return self.t.cat([self.op.apply(V[:, c0:c0 + 128].contiguous()) for c0 in range(0, V.shape[1], 128)], dim=1)

    def cat(self, mats):
        return self.t.cat(mats, dim=1).contiguous()

    def gram(self, S):
        return dev.to_host(K.gram(S))                        # S^T S on the FP64 tensor cores -> small host matrix

    def gemm(self, A, B):
        Bd = dev.to_device(np.ascontiguousarray(B, dtype=np.float64))
        Y, _ = K.ttm_dense(A, tuple(A.shape), Bd, 1, True)   # Y[i, j] = sum_k A[i, k] B[k, j]
        return Y

    def small_eigh(self, H, k):
        """Top-k eigenpairs of a small symmetric host matrix through the cluster eigensolver, DESCENDING."""
        U, W, _ = K.symeig_top(dev.to_device(np.ascontiguousarray(0.5 * (H + H.T))), k, flipsign=False, want_values=True)
        return dev.to_host(W), dev.to_host(U)

    def cols(self, A, k):
        return A[:, :k].contiguous()

    def to_host(self, A):
        return dev.to_host(A)


def _orthonormalise(ops, S):
    """Orthonormal basis of span(S) through the Gram matrix (scaled columns, tiny directions dropped), or None."""
    G = ops.gram(S)
    d = 1.0 / np.sqrt(np.maximum(np.diag(G), 1e-300))
    w, V = ops.small_eigh(G * d[:, None] * d[None, :], G.shape[0])
    keep = w > w.max() * 1e-12
    if not keep.any() or not np.isfinite(w).all():
        return None
    return ops.gemm(S, (d[:, None] * V[:, keep]) / np.sqrt(w[keep])[None, :])


def block_krylov_top(ops, n, k, depth=6, tol=1e-11, max_iter=300, seed=12345):
    """Leading k eigenpairs of the symmetric PSD operator behind ``ops.apply``: thick-restarted block Lanczos with full
    re-orthogonalisation.  Each sweep extends the block of Ritz vectors X by ``depth - 1`` Krylov blocks
    (A X, A^2 X, ... made orthonormal against everything before them), solves the projected eigenproblem and restarts
    from the best ``m = k + guards`` Ritz vectors.  Every tall operation is an operator application, a Gram or a GEMM;
    the host only sees the small coefficient matrices.  Returns (eigenvalues descending, n x k orthonormal block,
    sweeps, largest residual norm)."""
    m = min(n, k + max(2, k // 2))
    depth = max(2, min(depth, 512 // (2 * m)))               # the Gram of [Q, AQ] is at most 512 wide
    X = _orthonormalise(ops, ops.random(n, m, seed))
    AX = ops.apply(X)
    wl, rn, it = None, np.array([np.inf]), 0
    for it in range(max_iter):
        blocks, ablocks = [X], [AX]
        for _ in range(1, depth):
            Qall = ops.cat(blocks)
            q = Qall.shape[1]
            N = ablocks[-1]
            for _rep in range(2):                            # N -= Q (Q^T N), twice
                Cq = ops.gram(ops.cat([Qall, N]))[:q, q:]
                N = ops.gemm(ops.cat([N, Qall]), np.vstack([np.eye(N.shape[1]), -Cq]))
            N = _orthonormalise(ops, N)
            if N is None:                                    # the Krylov space is exhausted (low-rank operator)
                break
            blocks.append(N)
            ablocks.append(ops.apply(N))
        Qall, AQ = ops.cat(blocks), ops.cat(ablocks)
        q = Qall.shape[1]
        H = ops.gram(ops.cat([Qall, AQ]))[:q, q:]
        mm = min(m, q)
        wl, Y = ops.small_eigh(H, mm)
        X, AX = ops.gemm(Qall, Y), ops.gemm(AQ, Y)
        kk = min(k, mm)
        E = np.eye(mm)[:, :kk]
        R = ops.gemm(ops.cat([AX, X]), np.vstack([E, -E * wl[None, :kk]]))       # A X - X Lambda, no difference of squares
        rn = np.sqrt(np.maximum(np.diag(ops.gram(R)), 0.0))
        if rn.max() <= tol * max(abs(wl[0]), 1e-300) or len(blocks) == 1:
            break
    return wl[:k], ops.cols(X, k), it + 1, float(rn.max())


def _dense_gram_nvecs(op, rank, do_flipsign):
    t = dev.torch()
    n = op.rows
    cols = []
    for c0 in range(0, n, _BLOCK):
        b = min(_BLOCK, n - c0)
        E = dev.zeros((n, b))
        E[c0:c0 + b] = t.eye(b, dtype=t.float64, device=E.device)
        cols.append(op.apply(E))
    G = t.cat(cols, dim=1).contiguous()
    from .core import device_nvecs
    return dev.to_host(device_nvecs(G, rank, do_flipsign))


def _lobpcg_nvecs(op, rank, do_flipsign):
    ops = _DeviceOps(op)
    _, Xk, _, _ = block_krylov_top(ops, op.rows, int(rank))
    U = np.array(ops.to_host(Xk), dtype=float)
    if do_flipsign:
        rows = np.abs(U).argmax(axis=0)
        neg = U[rows, np.arange(U.shape[1])] < 0
        U[:, neg] = -U[:, neg]
    return U


def sparse_nvecs(X, n, rank, do_flipsign=True):
    """Host ndarray (I_n x rank): leading eigenvectors of the mode-n Gram of the sparse tensor X, descending."""
    dev.require_cuda()
    op = SparseGramOperator(X, n)
    try:
        path = os.environ.get('SKT_SPARSE_NVECS', '')
        if path == 'gram' or (path != 'lobpcg' and op.rows <= _DENSE_MAX):
            return np.array(_dense_gram_nvecs(op, rank, do_flipsign), dtype=float)
        return _lobpcg_nvecs(op, rank, do_flipsign)
    finally:
        op.close()

"""RESCAL-ALS -- drop-in for ``sktensor/rescal.py`` (reference: rescal.py:41-208, updates :212-265).

Same signature, defaults (``maxIter=100``, ``conv=1e-4``, ``init='nvecs'``), return tuple and quirks as the reference:
the third return value is the reference's never-assigned ``f = 0`` (rescal.py:170,208); the objective that steers the
convergence test is returned nowhere, exactly as there.

Everything that scales with ``n`` or the number of non-zeros runs on the device, on the kernels of the CP-ALS path:

* ``X_k B`` and ``X_k^T B`` (sparse slice times tall matrix): the sparse MTTKRP kernel on the 2-way view of the slice
  (``skt_mttkrp_coo``: all columns in one pass, segmented reduction, no atomics);
* ``A^T A``, ``A^T (X_k A)``: one tall Gram on the FP64 tensor cores (``skt_gram`` of ``[A, X_k A]``);
* ``A R_k``, ``F (lambda I + E)^-1``: tall-times-small GEMMs (``skt_ttm_dense`` on the 2-way view).

The ``rank x rank`` coefficient algebra between those calls (the reference's ``kron`` / ``solve`` / ``svd`` lines) stays
in NumPy on the host: it is O(k rank^3) and never touches data of size ``n``.  ``_updateR`` is evaluated without the
thin SVD of ``A``: with ``A^T A = V S^2 V^T``, ``Shat * (U^T X_k U) = (V^T A^T X_k A V) / (s_i^2 s_j^2 + lambda_R)``
(rescal.py:235-246), so the left singular vectors are never formed.
"""
import logging
import time

import numpy as np

__all__ = ['als', 'sptensor_to_list']

_DEF_MAXITER = 100
_DEF_INIT = 'nvecs'
_DEF_CONV = 1e-4
_DEF_LMBDA = 0
_DEF_ATTR = []
_DEF_NO_FIT = 1e9
_DEF_FIT_METHOD = None

_log = logging.getLogger('RESCAL')


class _Slice(object):
    """One frontal slice (or attribute matrix) on the device: ``dot(B) = M B``, ``tdot(B) = M^T B``."""

    def __init__(self, M):
        from scipy.sparse import issparse
        from . import _device as dev
        from . import _kernels as K
        self.shape = tuple(int(s) for s in M.shape)
        if issparse(M):
            C = M.tocoo()
            rows, cols, vals = C.row, C.col, C.data
        else:
            D = np.asarray(M, dtype=np.float64)
            rows, cols = np.nonzero(D)
            vals = D[rows, cols]
        self.normsq = float(np.sum(np.asarray(vals, dtype=np.float64) ** 2))
        self.nnz = len(vals)
        self.coo = None
        if self.nnz:
            self.coo = K.CooHandle([np.asarray(rows, dtype=np.int64), np.asarray(cols, dtype=np.int64)],
                                   np.asarray(vals, dtype=np.float64), self.shape)
        self.dev, self.K = dev, K

    def _apply(self, B, mode):
        rows = self.shape[mode]
        b = int(B.shape[1])
        if self.coo is None:
            return self.dev.zeros((rows, b))
        t = self.dev.torch()
        outs = []
        for c0 in range(0, b, 128):                          # the sparse kernel holds at most 128 columns per pass
            Bc = B[:, c0:c0 + 128].contiguous() if b > 128 else B
            U = [None, Bc] if mode == 0 else [Bc, None]
            outs.append(self.K.mttkrp_coo(self.coo, U, int(Bc.shape[1]), mode))
        return outs[0] if len(outs) == 1 else t.cat(outs, dim=1)

    def dot(self, B):
        return self._apply(B, 0)

    def tdot(self, B):
        return self._apply(B, 1)

    def close(self):
        if self.coo is not None:
            self.coo.close()
            self.coo = None


def _gemm(dev, K, A, B):
    """Device (rows x m) times host (m x p) -> device (rows x p)."""
    Y, _ = K.ttm_dense(A, tuple(A.shape), dev.to_device(np.ascontiguousarray(B, dtype=np.float64)), 1, True)
    return Y


def _cross(dev, K, A, B):
    """Host copies of ``A^T A`` and ``A^T B`` from one Gram of ``[A, B]``."""
    t = dev.torch()
    r = int(A.shape[1])
    G = dev.to_host(K.gram(t.cat([A, B], dim=1).contiguous()))
    return G[:r, :r], G[:r, r:]


def als(X, rank, **kwargs):
    """RESCAL-ALS: ``X_k ~ A R_k A^T`` for a list of ``n x n`` frontal slices (SciPy sparse matrices or arrays).

    Parameters (rescal.py:55-85): ``lambda_A``, ``lambda_R``, ``lambda_V`` (regularisation, default 0), ``attr`` (list of
    ``n x L_l`` attribute matrices), ``init`` ('nvecs' default, 'random'), ``compute_fval``, ``maxIter`` (100),
    ``conv`` (1e-4), ``orthogonalize``, ``dtype``.  Returns ``(A, R, f, itr, exectimes)``.
    """
    ainit = kwargs.pop('init', _DEF_INIT)
    maxIter = kwargs.pop('maxIter', _DEF_MAXITER)
    conv = kwargs.pop('conv', _DEF_CONV)
    lmbdaA = kwargs.pop('lambda_A', _DEF_LMBDA)
    lmbdaR = kwargs.pop('lambda_R', _DEF_LMBDA)
    lmbdaV = kwargs.pop('lambda_V', _DEF_LMBDA)
    func_compute_fval = kwargs.pop('compute_fval', _DEF_FIT_METHOD)
    orthogonalize = kwargs.pop('orthogonalize', False)
    P = kwargs.pop('attr', _DEF_ATTR)
    kwargs.pop('dtype', float)
    if not len(kwargs) == 0:
        raise ValueError('Unknown keywords (%s)' % (kwargs.keys()))

    sz = X[0].shape
    for Xi in X:
        if Xi.ndim != 2:
            raise ValueError('Frontal slices of X must be matrices')
        if Xi.shape != sz:
            raise ValueError('Frontal slices of X must be all of same shape')

    # which objective steers the convergence test (rescal.py:143-151); callables are honoured as in the reference
    if func_compute_fval is None:
        if orthogonalize:
            fval_kind = 'orth'
        elif np.prod(X[0].shape) * len(X) > _DEF_NO_FIT:
            _log.warning('For large tensors automatic computation of fit is disabled by default\n'
                         'To compute the fit, call rescal.als with "compute_fit=True"\n'
                         'Please note that this might cause memory and runtime problems')
            fval_kind = None
        else:
            fval_kind = 'full'
    else:
        fval_kind = func_compute_fval

    from . import _device as dev
    from . import _kernels as K
    dev.require_cuda()
    n, k = int(sz[0]), len(X)
    rank = int(rank)
    slices = [_Slice(Xi) for Xi in X]
    attrs = [_Slice(Pi) for Pi in P]
    try:
        if ainit == 'random':
            A = dev.to_device(np.array(np.random.rand(n, rank), dtype=float))
        elif ainit == 'nvecs':
            A = dev.to_device(_init_nvecs(dev, K, slices, n, rank))
        else:
            raise ValueError('Unknown init option ("%s")' % ainit)

        R, XA_AtA = _update_r(dev, K, slices, A, lmbdaR)
        Z = _update_z(dev, K, A, attrs, lmbdaV)
        fit = fitold = f = 0
        exectimes = []
        itr = -1
        for itr in range(maxIter):
            tic = time.time()
            fitold = fit
            A = _update_a(dev, K, slices, A, R, attrs, Z, lmbdaA, orthogonalize)
            R, (AtA, AtXA) = _update_r(dev, K, slices, A, lmbdaR)
            Z = _update_z(dev, K, A, attrs, lmbdaV)
            if callable(fval_kind):
                fit = fval_kind(X, dev.to_host(A), R, P, [dev.to_host(z) for z in Z], lmbdaA, lmbdaR, lmbdaV,
                                [s.normsq for s in slices])
            elif fval_kind is not None:
                fit = _fval(fval_kind, slices, AtA, AtXA, R, lmbdaA, lmbdaR)
            else:
                fit = np.inf
            fitchange = abs(fitold - fit)
            exectimes.append(time.time() - tic)
            _log.debug('[%3d] fval: %0.5f | delta: %7.1e | secs: %.5f' % (itr, fit, fitchange, exectimes[-1]))
            if itr > 0 and fitchange < conv:
                break
        return np.array(dev.to_host(A), dtype=float), R, f, itr + 1, np.array(exectimes)
    finally:
        for s in slices + attrs:
            s.close()


def _init_nvecs(dev, K, slices, n, rank):
    """``eigsh(sum_k X_k + X_k^T, rank)`` (rescal.py:171-177): the ``rank`` eigenvalues of LARGEST MAGNITUDE of the
    symmetrised sum, in ascending algebraic order like ARPACK returns them.  They are the leading eigenvectors of
    ``S^2`` (positive semi-definite); S is applied through the sparse kernel, never formed for long modes."""
    from ._eigs import _DeviceOps, block_krylov_top
    t = dev.torch()

    class _SymSum(object):
        rows = n

        def apply(self, V):
            out = None
            for s in slices:
                y = s.dot(V) + s.tdot(V)
                out = y if out is None else out + y
            return out

    op = _SymSum()
    if n <= K.SYMEIG_MAX_N:
        cols = []
        for c0 in range(0, n, 64):
            b = min(64, n - c0)
            E = dev.zeros((n, b))
            E[c0:c0 + b] = t.eye(b, dtype=t.float64, device=E.device)
            cols.append(op.apply(E))
        S = t.cat(cols, dim=1).contiguous()
        S = (0.5 * (S + S.t())).contiguous()
        kk = min(rank, n)
        Up, Wp, _ = K.symeig_top(S, kk, flipsign=False, want_values=True)
        Um, Wm, _ = K.symeig_top((-S).contiguous(), kk, flipsign=False, want_values=True)
        w = np.concatenate([dev.to_host(Wp), -dev.to_host(Wm)])
        V = np.concatenate([dev.to_host(Up), dev.to_host(Um)], axis=1)
    else:
        class _Sq(object):
            rows = n

            def apply(self, V):
                return op.apply(op.apply(V))
        ops = _DeviceOps(_Sq())
        _, Xk, _, _ = block_krylov_top(ops, n, rank)
        # Rayleigh quotients with S itself give the signed eigenvalues
        SX = op.apply(Xk)
        G = dev.to_host(K.gram(t.cat([Xk, SX], dim=1).contiguous()))
        w = np.diag(G[:rank, rank:]).copy()
        V = dev.to_host(Xk)
    pick = np.argsort(-np.abs(w), kind='stable')[:rank]
    pick = pick[np.argsort(w[pick], kind='stable')]
    return np.array(V[:, pick], dtype=float)


def _update_r(dev, K, slices, A, lmbdaR):
    """rescal.py:235-246 without the thin SVD (see the module docstring).  Returns R and the r x r pieces the objective
    re-uses: ``A^T A`` and the list of ``A^T X_k A``."""
    AtA = None
    AtXA = []
    for s in slices:
        g, c = _cross(dev, K, A, s.dot(A))
        AtA = g if AtA is None else AtA
        AtXA.append(c)
    if AtA is None:
        AtA = dev.to_host(K.gram(A))
    s2, V = np.linalg.eigh(0.5 * (AtA + AtA.T))
    s2 = np.maximum(s2, 0.0)
    den = np.outer(s2, s2) + lmbdaR
    with np.errstate(divide='ignore', invalid='ignore'):
        R = [V.dot(V.T.dot(C).dot(V) / den).dot(V.T) for C in AtXA]
    return R, (AtA, AtXA)


def _update_z(dev, K, A, attrs, lmbdaV):
    """rescal.py:249-262: ``Z_l = (A^T A + lambda I)^-1 A^T P_l`` (rank x L_l), kept on the device."""
    if not attrs:
        return []
    AtA = dev.to_host(K.gram(A))
    inv = np.linalg.inv(AtA + lmbdaV * np.eye(AtA.shape[0]))
    Z = []
    for p_ in attrs:
        PtA = p_.tdot(A)                                    # (L x rank)
        Z.append(_gemm(dev, K, PtA, inv.T).t().contiguous())  # (rank x L)
    return Z


def _update_a(dev, K, slices, A, R, attrs, Z, lmbdaA, orthogonalize):
    """rescal.py:212-232."""
    rank = int(A.shape[1])
    AtA = dev.to_host(K.gram(A))
    F = None
    E = np.zeros((rank, rank))
    for s, Rk in zip(slices, R):
        y = s.dot(_gemm(dev, K, A, Rk.T)) + s.tdot(_gemm(dev, K, A, Rk))
        F = y if F is None else F + y
        E += Rk.dot(AtA).dot(Rk.T) + Rk.T.dot(AtA).dot(Rk)
    for p_, Zl in zip(attrs, Z):
        y = p_.dot(Zl.t().contiguous())
        F = y if F is None else F + y
        E += dev.to_host(K.gram(Zl.t().contiguous()))        # Z Z^T
    M = np.linalg.inv(lmbdaA * np.eye(rank) + E)             # A = F (lambda I + E)^-1  (E is symmetric)
    Anew = _gemm(dev, K, F, M)
    if orthogonalize:                                        # orth(A) = U V^T of the thin SVD = A (A^T A)^-1/2
        G = dev.to_host(K.gram(Anew))
        w, V = np.linalg.eigh(0.5 * (G + G.T))
        Anew = _gemm(dev, K, Anew, (V / np.sqrt(np.maximum(w, 1e-300))).dot(V.T))
    return Anew


def _fval(kind, slices, AtA, AtXA, R, lmbdaA, lmbdaR):
    """rescal.py:265-278 from r x r quantities: ``||X_k - A R_k A^T||^2 = ||X_k||^2 - 2 <R_k, A^T X_k A> +
    tr(R_k^T A^T A R_k A^T A)``, and ``||A||^2 = tr(A^T A)``."""
    f = lmbdaA * float(np.trace(AtA))
    for s, C, Rk in zip(slices, AtXA, R):
        nr2 = float(np.sum(Rk * Rk))
        if kind == 'orth':
            f += (s.normsq - nr2) / s.normsq + lmbdaR * nr2
        else:
            res = s.normsq - 2.0 * float(np.sum(Rk * C)) + float(np.sum(Rk.T.dot(AtA).dot(Rk) * AtA))
            f += res / s.normsq + lmbdaR * nr2
    return f


def sptensor_to_list(X):
    """Frontal slices of a third-order ``sptensor`` as SciPy sparse matrices (rescal.py:281-294)."""
    from scipy.sparse import lil_matrix
    if X.ndim != 3:
        raise ValueError('Only third-order tensors are supported (ndim=%d)' % X.ndim)
    if X.shape[0] != X.shape[1]:
        raise ValueError('First and second mode must be of identical length')
    N, Kk = X.shape[0], X.shape[2]
    res = [lil_matrix((N, N)) for _ in range(Kk)]
    for i in range(len(X.vals)):
        res[X.subs[2][i]][X.subs[0][i], X.subs[1][i]] = X.vals[i]
    return res

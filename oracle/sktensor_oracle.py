"""CPU oracle for the CP-ALS / MTTKRP / HOOI hot path of scikit-tensor.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it; the shipped ``sktensor`` package never
does (it fails loudly when the CUDA library is missing).

It is a flat, functional NumPy restatement of the reference's algorithm; every
function cites the reference ``file:line`` (relative to the upstream
scikit-tensor tree) it follows, and keeps the reference's *operation order*
(full-tensor copies, per-rank-column passes, ``np.kron`` Khatri-Rao, SVD
``pinv`` ...) so that timing it is a fair stand-in for timing the reference on
the same host.

Parity status: PINNED.  ``oracle/make_golden.py`` runs the unmodified reference
(under the alias-only shim ``oracle/ref_shim.py``) in this container and writes
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function
here against those vectors and against the reference's own known-answer tests
(``sktensor/tests/test_base.py:17-81``, ``test_sptensor.py:121-149``).
"""
import time

import numpy as np
from scipy.linalg import eigh as _eigh
from scipy.linalg import pinv as _pinv


# ----------------------------------------------------------------------------
# index bookkeeping
# ----------------------------------------------------------------------------
def modes_without(ndim, skip):
    """All modes but ``skip`` ascending (pyutils.py:49-62 ``from_to_without``)."""
    return [m for m in range(ndim) if m != skip]


# ----------------------------------------------------------------------------
# dense tensor primitives
# ----------------------------------------------------------------------------
def unfold(X, mode):
    """Mode-``mode`` matricisation, Kolda-Bader column order (dtensor.py:103-150).

    Axis order is ``[mode, N-1, ..., 0 (mode skipped)]`` followed by a C-order
    reshape, so the lowest remaining mode varies fastest along the columns.  The
    reference pays two full copies here (dtensor.py:171 ``array(self)`` and the
    non-contiguous reshape at dtensor.py:149); so does this restatement.
    """
    X = np.asarray(X)
    N = X.ndim
    rest = [m for m in range(N - 1, -1, -1) if m != mode]
    moved = np.transpose(np.array(X), axes=[mode] + rest)
    ncols = 1
    for m in rest:
        ncols *= X.shape[m]
    return moved.reshape((X.shape[mode], ncols))


def fold(Xn, mode, shape):
    """Inverse of :func:`unfold` (dtensor.py:188-194)."""
    N = len(shape)
    rest = [m for m in range(N - 1, -1, -1) if m != mode]
    order = [mode] + rest
    arr = np.asarray(Xn).reshape([shape[m] for m in order])
    return np.transpose(arr, np.argsort(order))


def khatrirao(mats, reverse=False):
    """Column-wise Khatri-Rao product (core.py:333-378).

    Built one rank column at a time with a chain of ``np.kron`` exactly as the
    reference does (core.py:372-377); ``reverse=True`` walks the matrices last
    to first, which makes the FIRST listed matrix the fastest-varying one.
    """
    if not isinstance(mats, tuple):
        raise ValueError('A must be a tuple of array likes')
    ncol = mats[0].shape[1]
    nrow = 1
    for i, A in enumerate(mats):
        if A.ndim != 2:
            raise ValueError('A must be a tuple of matrices (A[%d].ndim = %d)' % (i, A.ndim))
        if A.shape[1] != ncol:
            raise ValueError('All matrices must have same number of columns')
        nrow *= A.shape[0]
    order = list(range(len(mats)))
    if reverse:
        order.reverse()
    out = np.zeros((nrow, ncol), dtype=mats[0].dtype)
    for r in range(ncol):
        col = mats[order[0]][:, r]
        for j in order[1:]:
            col = np.kron(col, mats[j][:, r])
        out[:, r] = col
    return out


def mttkrp_dense(X, U, mode):
    """Dense MTTKRP ``X_(mode) . (U_{N-1} kr ... kr U_0 without mode)`` (dtensor.py:163-167).

    ``U[mode]`` is ignored and may be ``None`` (cp.py:194,143).
    """
    X = np.asarray(X)
    others = modes_without(X.ndim, mode)
    Z = khatrirao(tuple(U[m] for m in others), reverse=True)
    return unfold(X, mode).dot(Z)


def norm_dense(X):
    """Frobenius norm of a dense tensor (dtensor.py:152-161)."""
    return np.linalg.norm(np.asarray(X))


def ttv_dense(X, vecs, dims):
    """Dense tensor-times-vectors over ``dims`` (core.py:105-127 + dtensor.py:78-98).

    ``vecs[i]`` multiplies mode ``dims[i]`` (ascending).  A permuting full copy
    (dtensor.py:90 -> :171) is followed by one reshape+``dot`` per contracted mode,
    trailing mode first.
    """
    X = np.asarray(X)
    for v, d in zip(vecs, dims):
        if len(v) != X.shape[d]:
            raise ValueError('Multiplicant is wrong size')
    rem = [m for m in range(X.ndim) if m not in dims]
    order = rem + list(dims)
    T = np.transpose(np.array(X), order)
    sz = np.array(X.shape)[order]
    nd = X.ndim
    for i in range(len(dims), 0, -1):
        T = T.reshape((int(sz[:nd - 1].prod()), int(sz[nd - 1])))
        T = T.dot(vecs[i - 1])
        nd -= 1
    if nd > 0:
        T = T.reshape(sz[:nd])
    return T


def ttm_dense(X, V, mode, transp=False):
    """Mode-``mode`` tensor-times-matrix (dtensor.py:58-76).

    ``transp=False``: ``Y = X x_mode V`` (V is p x I_mode); ``transp=True``: ``V`` is
    I_mode x p and is applied transposed.
    """
    X = np.asarray(X)
    sz = list(X.shape)
    order = [mode] + [m for m in range(X.ndim) if m != mode]
    mat = np.transpose(np.array(X), order).reshape(sz[mode], -1)
    if transp:
        mat = V.T.dot(mat)
        p = V.shape[1]
    else:
        mat = V.dot(mat)
        p = V.shape[0]
    out = mat.reshape([p] + sz[:mode] + sz[mode + 1:])
    return np.transpose(out, np.argsort(order))


def ttm_chain(X, U, skip_mode=None, transp=True):
    """TTM with every matrix of ``U`` but ``skip_mode``, ascending modes (core.py:50-103).

    This is ``ttm(X, U, n, transp=True, without=True)`` of tucker.py:103 (and, with
    ``skip_mode=None``, the all-modes product of tucker.py:133).
    """
    Y = np.asarray(X)
    for d in range(Y.ndim):
        if d == skip_mode:
            continue
        Y = ttm_dense(Y, U[d], d, transp)
    return Y


def flipsign(U):
    """Make each column's largest-magnitude entry positive (core.py:297-306)."""
    pos = np.abs(U).argmax(axis=0)
    for c in range(U.shape[1]):
        if U[pos[c], c] < 0:
            U[:, c] = -U[:, c]
    return U


def nvecs_dense(X, mode, rank):
    """Leading ``rank`` eigenvectors of ``X_(mode) X_(mode)^T`` (core.py:275-294).

    Descending eigenvalue order, sign-fixed with :func:`flipsign`.
    """
    Xn = unfold(X, mode)
    Y = Xn.dot(Xn.T)
    n = Y.shape[0]
    rank = int(rank)
    _, V = _eigh(Y, subset_by_index=(n - rank, n - 1))
    V = np.array(V[:, ::-1])
    return flipsign(V)


# ----------------------------------------------------------------------------
# sparse (COO) primitives -- a tensor is the triple (subs, vals, shape)
# ----------------------------------------------------------------------------
def coo_check(subs, vals):
    """Constructor checks of sptensor.py:68-77."""
    if not isinstance(subs, tuple):
        raise ValueError('Subscripts must be a tuple of array-likes')
    if len(subs[0]) != len(vals):
        raise ValueError('Subscripts and values must be of equal length')
    for s in subs:
        if np.array(s).dtype.kind != 'i':
            raise ValueError('Subscripts must be integers')


def norm_coo(vals):
    """Frobenius norm of a COO tensor; duplicates are NOT merged (sptensor.py:274-282)."""
    return np.linalg.norm(np.asarray(vals))


def coo_fromarray(A):
    """Dense -> COO (sptensor.py:362-366)."""
    A = np.asarray(A)
    subs = np.nonzero(A)
    return subs, A[subs], A.shape


def coo_toarray(subs, vals, shape):
    """COO -> dense; the last duplicate wins (sptensor.py:284-287)."""
    A = np.zeros(shape)
    A.put(np.ravel_multi_index(subs, tuple(shape)), vals)
    return A


def ttv_coo(subs, vals, shape, vecs, dims):
    """Sparse tensor-times-vectors (sptensor.py:153-177).

    Returns a scalar when every mode is contracted (case 1, :161-163) and, when one
    mode remains (case 2, :166-174), the pair ``(rows, sums)`` of the surviving
    1-D sparse vector with exact zeros dropped.
    """
    w = np.asarray(vals)
    for v, d in zip(vecs, dims):
        if len(v) != shape[d]:
            raise ValueError('Multiplicant is wrong size')
        w = w * v[subs[d]]
    rem = [m for m in range(len(shape)) if m not in dims]
    if not rem:
        return w.sum()
    if len(rem) != 1:
        raise NotImplementedError('oracle covers the 0- and 1-remaining-mode cases of the path')
    keys = subs[rem[0]]
    rows = np.unique(keys)
    bins = rows.searchsorted(keys)
    sums = np.bincount(bins, weights=w)
    (nz,) = sums.nonzero()
    return rows[nz], sums[nz]


def mttkrp_coo(subs, vals, shape, U, mode):
    """Sparse MTTKRP, one ``ttv`` pass per rank column (sptensor.py:216-229).

    ``R`` is read from ``U[1]`` for mode 0 and from ``U[0]`` otherwise (:218); the
    result is always a float64 ``(I_mode, R)`` array with zero rows for absent
    indices (:221); duplicate coordinates are summed by ``bincount`` (:172).
    """
    R = U[1].shape[1] if mode == 0 else U[0].shape[1]
    dims = modes_without(len(shape), mode)
    V = np.zeros((shape[mode], R))
    for r in range(R):
        cols = tuple(U[m][:, r] for m in dims)
        rows, sums = ttv_coo(subs, vals, shape, cols, dims)
        V[rows, r] = sums
    return V


# ----------------------------------------------------------------------------
# Kruskal tensor pieces
# ----------------------------------------------------------------------------
def kruskal_norm(U, lmbda):
    """``||P||`` from the Hadamard product of the factor Grams (ktensor.py:113-126)."""
    coef = np.outer(lmbda, lmbda)
    for Um in U:
        coef = coef * np.dot(Um.T, Um)
    return np.sqrt(coef.sum())


def kruskal_innerprod_dense(U, lmbda, X):
    """``<P, X>`` for dense X: one full ``ttv`` per rank column (ktensor.py:128-150)."""
    N = len(U)
    res = 0
    for r in range(len(lmbda)):
        res += lmbda[r] * ttv_dense(X, tuple(Um[:, r] for Um in U), list(range(N)))
    return res


def kruskal_innerprod_coo(U, lmbda, subs, vals, shape):
    """``<P, X>`` for COO X (ktensor.py:128-150 -> sptensor.py:154-163)."""
    N = len(U)
    res = 0
    for r in range(len(lmbda)):
        res += lmbda[r] * ttv_coo(subs, vals, shape, tuple(Um[:, r] for Um in U), list(range(N)))
    return res


def kruskal_toarray(U, lmbda):
    """Dense reconstruction of a Kruskal tensor (ktensor.py:152-163)."""
    shape = tuple(Um.shape[0] for Um in U)
    return np.dot(lmbda, khatrirao(tuple(U)).T).reshape(shape)


# ----------------------------------------------------------------------------
# CP-ALS
# ----------------------------------------------------------------------------
def gram_hadamard(U, mode, rank, dtype=float):
    """``Y = hadamard_{i != mode} U_i^T U_i`` (cp.py:144-146)."""
    Y = np.ones((rank, rank), dtype=dtype)
    for i in modes_without(len(U), mode):
        Y = Y * np.dot(U[i].T, U[i])
    return Y


def solve_normalise(M, Y, itr):
    """``Unew = M pinv(Y)``, then the reference's column scaling (cp.py:147-154).

    Iteration 0: column 2-norms.  Later iterations: the column *signed* maximum,
    with values below one replaced by one.  Returns ``(U_n, lambda)``.
    """
    Unew = M.dot(_pinv(Y))
    if itr == 0:
        lmbda = np.sqrt((Unew ** 2).sum(axis=0))
    else:
        lmbda = Unew.max(axis=0)
        lmbda[lmbda < 1] = 1
    return Unew / lmbda, lmbda


def cp_init(init, shape, rank, nvecs_fn=None, dtype=float):
    """Initial factors (cp.py:190-205): ``U[0]`` stays ``None``; ``'random'`` draws
    ``rand(I_n, rank)`` for n = 1..N-1 in order from the global legacy RNG."""
    N = len(shape)
    U = [None for _ in range(N)]
    if isinstance(init, list):
        U = init
    elif init == 'random':
        for n in range(1, N):
            U[n] = np.array(np.random.rand(shape[n], rank), dtype=dtype)
    elif init == 'nvecs':
        for n in range(1, N):
            U[n] = np.array(nvecs_fn(n, rank), dtype=dtype)
    else:
        raise ValueError('Unknown option (init=%s)' % str(init))
    return U


def cp_als(X, rank, init='nvecs', max_iter=500, fit_method='full', conv=1e-5,
           dtype=float, trace=None):
    """CP-ALS driver (cp.py:46-171).

    ``X`` is a dense ndarray or a COO triple ``(subs, vals, shape)``.  Returns
    ``(U, lmbda, fit, itr, exectimes)`` -- the reference's ``ktensor P`` is the pair
    ``(U, lmbda)``.  ``trace`` (optional list) receives the fit of every iteration.
    """
    sparse = isinstance(X, tuple)
    if sparse:
        subs, vals, shape = X
        shape = tuple(int(s) for s in shape)
        normX = norm_coo(vals)
        mttkrp = lambda U, n: mttkrp_coo(subs, vals, shape, U, n)
        innerprod = lambda U, l: kruskal_innerprod_coo(U, l, subs, vals, shape)
        nv = lambda n, r: nvecs_dense(coo_toarray(subs, vals, shape), n, r)
    else:
        X = np.asarray(X)
        shape = X.shape
        normX = norm_dense(X)
        mttkrp = lambda U, n: mttkrp_dense(X, U, n)
        innerprod = lambda U, l: kruskal_innerprod_dense(U, l, X)
        nv = lambda n, r: nvecs_dense(X, n, r)
    N = len(shape)
    U = cp_init(init, shape, rank, nv, dtype)
    fit = 0
    exectimes = []
    lmbda = None
    for itr in range(max_iter):
        tic = time.perf_counter()
        fitold = fit
        for n in range(N):
            M = mttkrp(U, n)
            Y = gram_hadamard(U, n, rank, dtype)
            U[n], lmbda = solve_normalise(M, Y, itr)
        if fit_method == 'full':
            normres = normX ** 2 + kruskal_norm(U, lmbda) ** 2 - 2 * innerprod(U, lmbda)
            fit = 1 - (normres / normX ** 2)
        else:
            fit = itr
        if trace is not None:
            trace.append(float(np.ravel(fit)[0]))
        exectimes.append(time.perf_counter() - tic)
        if itr > 0 and abs(fitold - fit) < conv:
            break
    return U, lmbda, fit, itr, np.array(exectimes)


# ----------------------------------------------------------------------------
# Tucker HOOI
# ----------------------------------------------------------------------------
def hooi(X, rank, init='nvecs', maxIter=500, conv=1e-7, trace=None):
    """Higher-order orthogonal iteration (tucker.py:36-123); dense X only.

    Returns ``(core, U, fit, itr)``; the reference returns ``(core, U)``.
    """
    X = np.asarray(X)
    N = X.ndim
    if np.isscalar(rank):
        rank = [rank] * N
    rank = [int(r) for r in rank]
    normX = norm_dense(X)
    if isinstance(init, list):
        U = init
    elif init == 'random':
        U = [None] + [np.array(np.random.rand(X.shape[n], rank[n])) for n in range(1, N)]
    elif init == 'nvecs':
        U = [None] + [nvecs_dense(X, n, rank[n]) for n in range(1, N)]
    else:
        raise ValueError('Unknown option (init=%s)' % str(init))
    fit = 0
    core = None
    for itr in range(maxIter):
        fitold = fit
        for n in range(N):
            Utilde = ttm_chain(X, U, n, transp=True)
            U[n] = nvecs_dense(Utilde, n, rank[n])
        core = ttm_dense(Utilde, U[N - 1], N - 1, transp=True)
        normres = np.sqrt(normX ** 2 - norm_dense(core) ** 2)
        fit = 1 - (normres / normX)
        if trace is not None:
            trace.append(float(fit))
        if itr > 1 and abs(fitold - fit) < conv:
            break
    return core, U, fit, itr


# ----------------------------------------------------------------------------
# RESCAL-ALS (rescal.py:41-265); X is a list of (sparse or dense) n x n frontal slices
# ----------------------------------------------------------------------------
def rescal_update_a(X, A, R, lmbda_a):
    """A <- [sum_k X_k A R_k^T + X_k^T A R_k] [lambda I + sum_k R_k A^T A R_k^T + R_k^T A^T A R_k]^-1 (rescal.py:212-232)."""
    rank = A.shape[1]
    F = np.zeros(A.shape)
    E = np.zeros((rank, rank))
    AtA = A.T.dot(A)
    for Xk, Rk in zip(X, R):
        F += Xk.dot(A.dot(Rk.T)) + Xk.T.dot(A.dot(Rk))
        E += Rk.dot(AtA).dot(Rk.T) + Rk.T.dot(AtA).dot(Rk)
    return np.linalg.solve(lmbda_a * np.eye(rank) + E.T, F.T).T


def rescal_update_r(X, A, lmbda_r):
    """R_k from the thin SVD of A (rescal.py:235-246)."""
    rank = A.shape[1]
    U, S, Vt = np.linalg.svd(A, full_matrices=False)
    Shat = np.kron(S, S)
    Shat = (Shat / (Shat ** 2 + lmbda_r)).reshape(rank, rank)
    return [Vt.T.dot(Shat * U.T.dot(Xk.dot(U))).dot(Vt) for Xk in X]


def rescal_fval(X, A, R, lmbda_a, lmbda_r, normX):
    """Objective the reference monitors for convergence (rescal.py:265-271, `_compute_fval`)."""
    f = lmbda_a * np.linalg.norm(A) ** 2
    for Xk, Rk, nk in zip(X, R, normX):
        D = (Xk.toarray() if hasattr(Xk, 'toarray') else np.asarray(Xk)) - A.dot(Rk).dot(A.T)
        f += np.linalg.norm(D) ** 2 / nk + lmbda_r * np.linalg.norm(Rk) ** 2
    return f


def rescal_als(X, rank, A0, maxIter=100, conv=1e-4, lambda_A=0.0, lambda_R=0.0, trace=None):
    """RESCAL-ALS without attributes, from the initial factor ``A0`` (rescal.py:166-208; the reference draws A0 with
    ``numpy.random.rand`` or from ``eigsh``).  Returns ``(A, R, fval_of_last_iteration, iterations)`` -- the reference's
    third return value is the never-assigned ``f = 0`` (rescal.py:170,208)."""
    A = np.array(A0, dtype=float)
    R = rescal_update_r(X, A, lambda_R)
    normX = [float(((Xk.data if hasattr(Xk, 'data') and not isinstance(Xk, np.ndarray) else np.asarray(Xk)) ** 2).sum()) for Xk in X]
    fit = 0
    itr = 0
    for itr in range(maxIter):
        fitold = fit
        A = rescal_update_a(X, A, R, lambda_A)
        R = rescal_update_r(X, A, lambda_R)
        fit = rescal_fval(X, A, R, lambda_A, lambda_R, normX)
        if trace is not None:
            trace.append(float(fit))
        if itr > 0 and abs(fitold - fit) < conv:
            break
    return A, R, fit, itr + 1

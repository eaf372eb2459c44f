"""NumPy model of ``symeig_cluster_kernel`` (scikit-tensor_b200/csrc/symeig.cu): the same four phases, the same
recurrences and guards, vectorised over the eigenvalue index where the kernel uses one thread per eigenvalue.

Development aid and executable documentation of the kernel's numerics (checked against LAPACK in
tests/test_symeig_model.py); the product never imports it.

  1. Householder tridiagonalisation  A = Q T Q^T          (kernel: matrix distributed by columns over a cluster)
  2. top-k eigenvalues of T by multisection on Sturm counts
  3. eigenvectors of T by inverse iteration on the pivoted LU of T - lambda I, modified Gram-Schmidt inside groups of
     close eigenvalues (the dstein recipe, all vectors advanced together)
  4. back-transformation  V = Q Z, descending order, sign fixed like sktensor.core.flipsign (core.py:297-306)
"""
import numpy as np

EPS = np.finfo(float).eps
SAFMIN = np.finfo(float).tiny


def tridiagonalise(A):
    """Lower Householder reduction.  Returns d, e, V (reflector k in V[k+1:, k], V[k+1, k] = 1), tau."""
    A = np.array(A, dtype=float)
    n = A.shape[0]
    d = np.zeros(n)
    e = np.zeros(max(n - 1, 0))
    tau = np.zeros(max(n - 1, 0))
    V = np.zeros((n, n))
    for k in range(n - 1):
        d[k] = A[k, k]
        x = A[k + 1:, k].copy()
        alpha = x[0]
        xnorm2 = float(np.dot(x[1:], x[1:]))
        if xnorm2 == 0.0:
            tau[k] = 0.0
            e[k] = alpha
            v = np.zeros_like(x)
            v[0] = 1.0
        else:
            beta = -np.copysign(np.sqrt(alpha * alpha + xnorm2), alpha)
            tau[k] = (beta - alpha) / beta
            v = x / (alpha - beta)
            v[0] = 1.0
            e[k] = beta
        V[k + 1:, k] = v
        if tau[k] != 0.0:
            B = A[k + 1:, k + 1:]
            p = tau[k] * B.dot(v)
            w = p - (0.5 * tau[k] * float(np.dot(p, v))) * v
            A[k + 1:, k + 1:] = B - np.outer(v, w) - np.outer(w, v)
    d[n - 1] = A[n - 1, n - 1]
    return d, e, V, tau


def sturm_count(d, e2, x, pivmin):
    """Number of eigenvalues of T strictly below each shift in ``x`` (vectorised over shifts)."""
    x = np.asarray(x, dtype=float)
    q = d[0] - x
    q = np.where(np.abs(q) < pivmin, -pivmin, q)
    cnt = (q < 0).astype(np.int64)
    for i in range(1, len(d)):
        q = d[i] - x - e2[i - 1] / q
        q = np.where(np.abs(q) < pivmin, -pivmin, q)
        cnt += (q < 0)
    return cnt


def top_eigenvalues(d, e, k, lanes=32, max_rounds=40):
    """The k largest eigenvalues of T, ascending, by multisection with ``lanes`` interior points per round."""
    n = len(d)
    e2 = e * e
    pivmin = SAFMIN * max(1.0, float(e2.max()) if n > 1 else 1.0)
    ae = np.concatenate(([0.0], np.abs(e), [0.0]))
    rad = ae[:-1] + ae[1:]
    gl, gu = float(np.min(d - rad)), float(np.max(d + rad))
    bnorm = max(abs(gl), abs(gu))
    gl -= 2.1 * bnorm * EPS * n + 2.1 * pivmin
    gu += 2.1 * bnorm * EPS * n + 2.1 * pivmin
    idx = np.arange(n - k, n)
    lo = np.full(k, gl)
    hi = np.full(k, gu)
    frac = (np.arange(lanes) + 1.0) / (lanes + 1.0)
    for _ in range(max_rounds):
        width = hi - lo
        if np.all(width <= 2.0 * EPS * np.maximum(np.abs(lo), np.abs(hi)) + 2.0 * pivmin):
            break
        xs = lo[:, None] + width[:, None] * frac[None, :]
        cnt = sturm_count(d, e2, xs.ravel(), pivmin).reshape(k, lanes)
        below = cnt <= idx[:, None]                       # eigenvalue idx lies at or above this shift
        nb = below.sum(axis=1)                            # counts are monotone in the shift
        new_lo = np.where(nb > 0, xs[np.arange(k), np.maximum(nb - 1, 0)], lo)
        new_hi = np.where(nb < lanes, xs[np.arange(k), np.minimum(nb, lanes - 1)], hi)
        lo, hi = np.maximum(lo, new_lo), np.minimum(hi, new_hi)
    return 0.5 * (lo + hi), bnorm


def lu_factor(d, e, lam):
    """dlagtf for every shift in ``lam`` at once: T - lam I = P L U.  Returns (a, b, c, d2, swp) with
    a = diag(U), b = first, d2 = second super-diagonal, c = multipliers, swp = interchange flags."""
    n, k = len(d), len(lam)
    a = d[:, None] - lam[None, :]
    b = np.repeat(e[:, None], k, axis=1) if n > 1 else np.zeros((0, k))
    c = b.copy()
    d2 = np.zeros((max(n - 2, 0), k))
    swp = np.zeros((max(n - 1, 0), k), dtype=bool)
    if n == 1:
        return a, b, c, d2, swp
    scale1 = np.abs(a[0]) + np.abs(b[0])
    for i in range(n - 1):
        scale2 = np.abs(c[i]) + np.abs(a[i + 1]) + (np.abs(b[i + 1]) if i < n - 2 else 0.0)
        with np.errstate(divide='ignore', invalid='ignore'):
            piv1 = np.where(a[i] == 0, 0.0, np.abs(a[i]) / scale1)
            piv2 = np.where(c[i] == 0, 0.0, np.abs(c[i]) / scale2)
            czero = c[i] == 0
            noswap = czero | (piv2 <= piv1)
            # no interchange
            m_ns = np.where(czero, 0.0, c[i] / np.where(a[i] == 0, 1.0, a[i]))
            a1_ns = a[i + 1] - m_ns * b[i]
            # interchange rows i, i+1
            m_sw = a[i] / np.where(czero, 1.0, c[i])
            a_sw = c[i].copy()
            a1_sw = b[i] - m_sw * a[i + 1]
            b_sw = a[i + 1].copy()
        if i < n - 2:
            d2[i] = np.where(noswap, 0.0, b[i + 1])
            b[i + 1] = np.where(noswap, b[i + 1], -m_sw * b[i + 1])
        a_i_new = np.where(noswap, a[i], a_sw)
        a[i + 1] = np.where(noswap, a1_ns, a1_sw)
        b[i] = np.where(noswap, b[i], b_sw)
        c[i] = np.where(noswap, m_ns, m_sw)
        a[i] = a_i_new
        swp[i] = ~noswap
        scale1 = scale2
    return a, b, c, d2, swp


def lu_solve(fac, y, tol):
    """dlagts(job = -1): solve (T - lam I) x = y in place; pivots below ``tol`` are replaced by +-tol."""
    a, b, c, d2, swp = fac
    n = a.shape[0]
    y = y.copy()
    for i in range(1, n):
        yi1, yi = y[i - 1].copy(), y[i].copy()
        y[i - 1] = np.where(swp[i - 1], yi, yi1)
        y[i] = np.where(swp[i - 1], yi1 - c[i - 1] * yi, yi - c[i - 1] * yi1)
    for i in range(n - 1, -1, -1):
        t = y[i].copy()
        if i <= n - 2:
            t = t - b[i] * y[i + 1]
        if i <= n - 3:
            t = t - d2[i] * y[i + 2]
        ak = np.where(np.abs(a[i]) < tol, np.where(a[i] < 0, -tol, tol), a[i])
        y[i] = t / ak
    return y


def lcg_uniform(n, k):
    """Deterministic start vectors in (-1, 1): the kernel hashes (row, column) the same way."""
    i = np.arange(n, dtype=np.uint64)[:, None]
    j = np.arange(k, dtype=np.uint64)[None, :]
    with np.errstate(over='ignore'):
        h = (i * np.uint64(0x9E3779B97F4A7C15) + j * np.uint64(0xC2B2AE3D27D4EB4F) + np.uint64(0x165667B19E3779F9))
        h ^= h >> np.uint64(29)
        h *= np.uint64(0xBF58476D1CE4E5B9)
        h ^= h >> np.uint64(32)
    return (h >> np.uint64(11)).astype(np.float64) * (2.0 / 9007199254740992.0) - 1.0


def tridiag_eigenvectors(d, e, lam, bnorm, iters=2):
    """Inverse iteration for all shifts together; modified Gram-Schmidt inside groups of close eigenvalues
    (gap < 1e-3 * ||T||, dstein's ortol) in ascending order after every solve."""
    n, k = len(d), len(lam)
    onenrm = max(bnorm, SAFMIN)
    ortol = 1e-3 * onenrm
    # dstein separates numerically equal shifts by 10 eps |lambda| so that the factorisations differ
    lam = lam.copy()
    for j in range(1, k):
        sep = 10.0 * EPS * abs(lam[j])
        if lam[j] - lam[j - 1] < sep:
            lam[j] = lam[j - 1] + sep
    group = np.zeros(k, dtype=np.int64)            # first index of the group each eigenvalue belongs to
    for j in range(1, k):
        group[j] = group[j - 1] if (lam[j] - lam[j - 1]) < ortol else j
    fac = lu_factor(d, e, lam)
    amax = np.maximum(np.abs(fac[0]).max(axis=0), np.abs(fac[1]).max(axis=0) if n > 1 else 0.0)
    if n > 2:
        amax = np.maximum(amax, np.abs(fac[3]).max(axis=0))
    tol = np.where(amax == 0, EPS, np.maximum(amax * EPS, SAFMIN / EPS))     # dlagts: tol = eps * max|U|, eps if U = 0
    Z = lcg_uniform(n, k)
    # sub-groups of numerically coincident eigenvalues: only these need orthogonalising BETWEEN the solves
    tight = np.zeros(k, dtype=np.int64)
    for j in range(1, k):
        tight[j] = tight[j - 1] if (lam[j] - lam[j - 1]) < 1e-8 * onenrm else j
    for it in range(iters):
        last = it == iters - 1
        # scale the right-hand side like dstein so that the solve cannot overflow
        nrm1 = np.abs(Z).sum(axis=0)
        scl = n * onenrm * np.maximum(EPS, np.abs(fac[0][n - 1])) / np.where(nrm1 == 0, 1.0, nrm1)
        Z = lu_solve(fac, Z * scl[None, :], tol)
        zmax = np.abs(Z).max(axis=0)                 # keep the squares of the next norms inside the double range
        Z = Z / np.where(zmax > 0, zmax, 1.0)[None, :]
        nz = np.linalg.norm(Z, axis=0)
        Z = Z / np.where(nz > 0, nz, 1.0)[None, :]
        grp = group if last else tight
        j0 = 0
        while j0 < k:
            j1 = j0 + 1
            while j1 < k and grp[j1] == grp[j0]:
                j1 += 1
            if j1 - j0 > 1:
                Zg = Z[:, j0:j1]
                E = Zg.T.dot(Zg) - np.eye(j1 - j0)
                if np.abs(E).max() < 1e-4 and j1 - j0 <= 32:
                    # nearly orthonormal already: symmetric (Loewdin) correction to third order, no sequential steps
                    Z[:, j0:j1] = Zg.dot(np.eye(j1 - j0) - 0.5 * E + 0.375 * E.dot(E))
                else:
                    for j in range(j0, j1):          # modified Gram-Schmidt, ascending
                        for i in range(j0, j):
                            Z[:, j] -= np.dot(Z[:, i], Z[:, j]) * Z[:, i]
                        nzj = np.linalg.norm(Z[:, j])
                        Z[:, j] /= nzj if nzj > 0 else 1.0
            j0 = j1
    return Z


def back_transform(V, tau, Z):
    X = Z.copy()
    n = V.shape[0]
    for k in range(n - 2, -1, -1):
        if tau[k] == 0.0:
            continue
        v = V[k + 1:, k]
        X[k + 1:] -= np.outer(v, tau[k] * v.dot(X[k + 1:]))
    return X


def symeig_top(A, k, iters=2, lanes=32):
    """(w, U): the k largest eigenvalues (descending) and eigenvectors of symmetric A, columns sign-fixed."""
    A = np.asarray(A, dtype=float)
    n = A.shape[0]
    d, e, V, tau = tridiagonalise(A)
    lam, bnorm = top_eigenvalues(d, e, k, lanes=lanes)
    Z = tridiag_eigenvectors(d, e, lam, bnorm, iters=iters)
    X = back_transform(V, tau, Z)
    X = X[:, ::-1]
    w = lam[::-1]
    for c in range(k):
        X[:, c] /= np.linalg.norm(X[:, c])
        r = int(np.argmax(np.abs(X[:, c])))
        if X[r, c] < 0:
            X[:, c] = -X[:, c]
    return w, X

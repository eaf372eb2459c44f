"""The restarted block Lanczos loop behind sparse nvecs (sktensor/_eigs.py: block_krylov_top), driven on CPU with NumPy
stand-ins for the device primitives (operator application, Gram, GEMM, small eigensolve) and compared with ARPACK --
what the reference calls for this step (core.py:280-283)."""
import importlib.util
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as sla

from conftest import PKG


class NumpyOps(object):
    def __init__(self, A):
        self.A, self.applied = A, 0

    def random(self, n, m, seed):
        return np.random.RandomState(seed).rand(n, m) - 0.5

    def apply(self, V):
        self.applied += V.shape[1]
        return self.A.dot(V)

    def cat(self, mats):
        return np.concatenate(mats, axis=1)

    def gram(self, S):
        return S.T.dot(S)

    def gemm(self, A, B):
        return A.dot(B)

    def small_eigh(self, H, k):
        w, V = np.linalg.eigh(0.5 * (H + H.T))
        return w[::-1][:k].copy(), V[:, ::-1][:, :k].copy()

    def cols(self, A, k):
        return A[:, :k]

    def to_host(self, A):
        return A


def _load_loop():
    # the loop is pure NumPy + the ops object; import it without the module's device imports
    src = open(os.path.join(PKG, 'sktensor', '_eigs.py')).read()
    start, end = src.index('def _orthonormalise'), src.index('def _dense_gram_nvecs')
    ns = {'np': np}
    exec(compile(src[start:end], '_eigs_loop', 'exec'), ns)
    return ns['block_krylov_top']


@pytest.mark.parametrize('n,cols,density,k', [(1500, 8000, 0.004, 4), (2500, 600, 0.01, 8), (400, 3000, 0.02, 16)])
def test_block_lanczos_matches_arpack(n, cols, density, k):
    loop = _load_loop()
    Xn = sp.random(n, cols, density=density, random_state=1, format='csr')
    A = (Xn @ Xn.T).tocsr()
    ops = NumpyOps(A)
    w, X, sweeps, res = loop(ops, n, k)
    wr, Vr = sla.eigsh(A, k, which='LM')
    wr, Vr = wr[::-1], Vr[:, ::-1]
    assert np.abs(w - wr).max() < 1e-12 * wr[0]
    assert np.abs(X.T.dot(X) - np.eye(k)).max() < 1e-12
    assert np.linalg.norm(X.dot(X.T) - Vr.dot(Vr.T)) < 1e-7
    assert res < 1e-9 * wr[0]


def test_block_lanczos_low_rank_operator():
    loop = _load_loop()
    rng = np.random.RandomState(0)
    B = rng.rand(300, 5)
    A = B.dot(B.T)                                   # rank 5 < k + guards: the Krylov space runs out
    w, X, _, _ = loop(NumpyOps(A), 300, 3)
    wr = np.linalg.eigvalsh(A)[::-1][:3]
    assert np.abs(w - wr).max() < 1e-10 * wr[0]

"""tools/symeig_model.py -- the NumPy model of the cluster eigensolver's numerics (tridiagonalisation, Sturm
multisection, pivoted-LU inverse iteration with in-group Gram-Schmidt, back-transformation) -- against LAPACK.
CPU only; the kernel itself is checked the same way in tests/test_gpu_symeig.py."""
import os
import sys

import numpy as np
import pytest
import scipy.linalg as sl

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, 'tools'))
import symeig_model as sm  # noqa: E402


def spectra():
    rng = np.random.RandomState(0)
    out = []
    for n, k in [(1, 1), (2, 1), (2, 2), (3, 2), (5, 5), (10, 3), (33, 7), (64, 64), (100, 10)]:
        B = rng.randn(n, n)
        out.append(('rand%d_%d' % (n, k), B + B.T, k, True))
    Y = rng.rand(160, 400)
    out.append(('gram_uniform', Y.dot(Y.T), 32, True))
    Q, _ = np.linalg.qr(rng.randn(60, 60))
    D = np.concatenate([5 + 1e-10 * rng.randn(10), 3 + 1e-7 * rng.randn(20), rng.rand(30)])
    out.append(('clustered_near', (Q * D).dot(Q.T), 30, True))
    D = np.array([5.0] * 10 + [3.0] * 20 + [1e-3] * 30)
    out.append(('clustered_exact', (Q * D).dot(Q.T), 15, False))      # the cut goes through a degenerate eigenspace
    out.append(('identity', np.eye(20), 5, False))
    out.append(('zero', np.zeros((12, 12)), 4, False))
    out.append(('diagonal', np.diag(np.arange(30.)), 6, True))
    Y = rng.rand(50, 8)
    out.append(('rank_deficient', Y.dot(Y.T), 20, False))
    n = 21
    W = np.diag(np.abs(np.arange(n) - 10.)) + np.diag(np.ones(n - 1), 1) + np.diag(np.ones(n - 1), -1)
    out.append(('wilkinson21', W, 6, True))
    return out


@pytest.mark.parametrize('name,A,k,unique', spectra(), ids=[s[0] for s in spectra()])
def test_model_matches_lapack(name, A, k, unique):
    n = A.shape[0]
    w, U = sm.symeig_top(A, k)
    wr, Vr = sl.eigh(A, subset_by_index=(n - k, n - 1))
    wr, Vr = wr[::-1], Vr[:, ::-1]
    scale = max(np.abs(sl.eigvalsh(A)).max(), 1e-300)
    assert np.abs(w - wr).max() <= 1e-13 * scale + 1e-300
    assert np.abs(U.T.dot(U) - np.eye(k)).max() < 1e-12
    assert np.linalg.norm(A.dot(U) - U * w[None, :], axis=0).max() <= 1e-13 * scale + 1e-300
    if unique:
        assert np.linalg.norm(U.dot(U.T) - Vr.dot(Vr.T)) < 1e-9
        # sign convention of core.flipsign: the largest-magnitude entry of every column is positive
        assert (U[np.abs(U).argmax(axis=0), np.arange(k)] > 0).all()

"""skt_symeig_top (csrc/symeig.cu) against LAPACK: the eigensolver behind core.nvecs (core.py:275-306) -- every size
class of the kernel's plan (1 / 4 / 8 / 16-CTA clusters, matrix in shared or global memory, factor arrays and the
orthogonalisation copy in shared or global memory), degenerate spectra, and the nvecs golden vectors."""
import numpy as np
import pytest
import scipy.linalg as sl

from conftest import relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def K():
    from sktensor import _device as dev, _kernels as K
    dev.require_cuda()
    return K


@pytest.fixture(scope='module')
def dev():
    from sktensor import _device as dev
    return dev


def run(K, dev, A, k, flip=True):
    U, W, info = K.symeig_top(dev.to_device(A), k, flipsign=flip, want_values=True)
    return dev.to_host(W), dev.to_host(U), int(dev.to_host(info)[0])


def check(A, k, w, U, unique=True, tol=1e-12):
    n = A.shape[0]
    wr, Vr = sl.eigh(A, subset_by_index=(n - k, n - 1))
    wr, Vr = wr[::-1], Vr[:, ::-1]
    scale = max(np.abs(sl.eigvalsh(A)).max(), 1e-300)
    assert np.abs(w - wr).max() <= tol * scale + 1e-300
    assert np.abs(U.T.dot(U) - np.eye(k)).max() < 1e-11
    assert np.linalg.norm(A.dot(U) - U * w[None, :], axis=0).max() <= tol * scale + 1e-300
    if unique:
        assert np.linalg.norm(U.dot(U.T) - Vr.dot(Vr.T)) < 1e-8
        assert (U[np.abs(U).argmax(axis=0), np.arange(k)] > 0).all()


@pytest.mark.parametrize('n,k', [(1, 1), (2, 1), (2, 2), (3, 2), (5, 5), (10, 3), (33, 7), (48, 48), (49, 5), (64, 64),
                                 (100, 10), (128, 128), (129, 17), (160, 64), (200, 200), (256, 32), (257, 32), (384, 32),
                                 (512, 32), (512, 512), (700, 16), (1024, 8)])
def test_random_symmetric_vs_lapack(K, dev, n, k):
    rng = np.random.RandomState(n * 1000 + k)
    B = rng.randn(n, n)
    A = B + B.T
    w, U, info = run(K, dev, A, k)
    assert info == 0
    check(A, k, w, U)


@pytest.mark.parametrize('n,cols,k', [(384, 1024, 32), (160, 4096, 64), (512, 512, 32), (96, 64, 8)])
def test_gram_of_uniform_data(K, dev, n, cols, k):
    """The matrices HOOI and init='nvecs' produce: one dominant eigenvalue (the mean) over a dense noise bulk."""
    rng = np.random.RandomState(n + cols)
    Y = rng.rand(n, cols)
    A = Y.dot(Y.T)
    w, U, info = run(K, dev, A, k)
    assert info == 0
    check(A, k, w, U, unique=(cols >= n))
    w2, U2, _ = run(K, dev, A, k)
    assert (w == w2).all() and (U == U2).all()                      # bitwise reproducible


def test_degenerate_spectra(K, dev):
    rng = np.random.RandomState(5)
    Q, _ = np.linalg.qr(rng.randn(60, 60))
    D = np.concatenate([5 + 1e-10 * rng.randn(10), 3 + 1e-7 * rng.randn(20), rng.rand(30)])
    A = (Q * D).dot(Q.T)
    A = 0.5 * (A + A.T)
    check(A, 30, *run(K, dev, A, 30)[:2])
    D = np.array([5.0] * 10 + [3.0] * 20 + [1e-3] * 30)
    A = (Q * D).dot(Q.T)
    A = 0.5 * (A + A.T)
    check(A, 15, *run(K, dev, A, 15)[:2], unique=False)
    check(np.eye(20), 5, *run(K, dev, np.eye(20), 5)[:2], unique=False)
    check(np.zeros((12, 12)), 4, *run(K, dev, np.zeros((12, 12)), 4)[:2], unique=False)
    Dg = np.diag(np.arange(30.))
    check(Dg, 6, *run(K, dev, Dg, 6)[:2])
    Y = rng.rand(50, 8)
    check(Y.dot(Y.T), 20, *run(K, dev, Y.dot(Y.T), 20)[:2], unique=False)
    n = 21
    W = np.diag(np.abs(np.arange(n) - 10.)) + np.diag(np.ones(n - 1), 1) + np.diag(np.ones(n - 1), -1)
    check(W, 6, *run(K, dev, W, 6)[:2])


def test_nonfinite_input_is_flagged(K, dev):
    A = np.eye(8)
    A[3, 3] = np.nan
    _, _, info = run(K, dev, A, 2)
    assert info == -1


def test_no_flipsign_and_nvecs_golden(K, dev, golden):
    from sktensor import dtensor
    from sktensor.core import nvecs
    rng = np.random.RandomState(1)
    B = rng.randn(40, 40)
    A = B + B.T
    _, U0, _ = run(K, dev, A, 5, flip=False)
    _, U1, _ = run(K, dev, A, 5, flip=True)
    assert relerr(np.abs(U0), np.abs(U1)) < 1e-14
    X = golden['cfg1_X']
    for n in range(3):
        got = nvecs(dtensor(X), n, 3)
        Xn = np.moveaxis(X, n, 0).reshape(X.shape[n], -1)
        w, V = np.linalg.eigh(Xn.dot(Xn.T))
        V = V[:, ::-1][:, :3]
        V = V * np.sign(V[np.abs(V).argmax(axis=0), np.arange(3)])
        assert relerr(got, V) < 1e-9

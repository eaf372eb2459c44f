"""The callers either side of the sparse path, on the device:
accumfun de-duplication (sptensor.py:80-84 / utils.py:5-26), sparse ttm (sptensor.py:139-151) and with it
tucker_hooi on an sptensor (tucker.py:103), sparse nvecs (core.py:280-283)."""
import numpy as np
import pytest

from conftest import relerr
from oracle import sktensor_oracle as orc

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


@pytest.mark.parametrize('shape,nnz,func', [((30, 20, 10), 20000, np.sum), ((1000, 50), 30000, np.max), ((7, 9, 11, 5), 9000, np.min),
                                            ((100000, 3000, 200), 50000, np.sum), ((50, 40, 30), 5000, lambda v: v[-1])])
def test_accumfun_matches_reference_accum(K, shape, nnz, func):
    from sktensor import sptensor
    from sktensor.utils import accum
    g = np.random.default_rng(nnz)
    subs = tuple(np.minimum((I * g.random(nnz) ** 2).astype(np.int64), I - 1) for I in shape)
    vals = g.random(nnz) - 0.5
    S = sptensor(subs, vals, shape=shape, accumfun=func)
    rv, rs = accum(subs, vals, func=func, with_subs=True)          # the reference's host loop (utils.py:5-26)
    assert len(S.vals) == len(rv) < nnz
    for a, b in zip(S.subs, rs):
        assert (np.asarray(a) == np.asarray(b)).all()              # same coordinates in the same (lexsort) order
    assert relerr(S.vals, rv) < 1e-13
    if func in (np.max, np.min):
        assert (np.asarray(S.vals) == rv).all()


def test_accumfun_negative_and_offset_subscripts(K):
    from sktensor.sptensor import _accumulate
    from sktensor.utils import accum
    g = np.random.default_rng(3)
    subs = (g.integers(-5, 6, 10000), g.integers(100, 140, 10000))
    vals = g.random(10000)
    v, s = _accumulate(subs, vals, np.sum)
    rv, rs = accum(subs, vals, func=np.sum, with_subs=True)
    assert all((np.asarray(a) == np.asarray(b)).all() for a, b in zip(s, rs)) and relerr(v, rv) < 1e-13


def test_compact_keys(K, dev):
    t = dev.torch()
    g = np.random.default_rng(0)
    keys = g.integers(0, 1 << 40, 5000, dtype=np.int64)[g.integers(0, 5000, 60000)]
    ids, nd = K.compact_keys(t.from_numpy(keys).cuda(), 41)       # non-negative int64 keys are valid uint64 keys
    u, inv = np.unique(keys, return_inverse=True)
    assert nd == len(u) and (dev.to_host(ids) == inv).all()


@pytest.mark.parametrize('shape,nnz,p', [((30, 20, 10), 2000, 4), ((12, 14, 16, 6), 3000, 7), ((200, 150), 5000, 33),
                                         ((25, 18, 20), 4000, 150), ((9, 8, 7), 1, 3)])
def test_sparse_ttm_matches_dense(K, shape, nnz, p):
    from sktensor import sptensor
    g = np.random.default_rng(nnz + p)
    subs = tuple(g.integers(0, I, nnz) for I in shape)
    vals = g.random(nnz) - 0.3
    S = sptensor(subs, vals, shape=shape, accumfun=np.sum)         # distinct coordinates (toarray keeps the last duplicate)
    A = S.toarray()
    for mode in range(len(shape)):
        for transp in (False, True):
            V = g.random((shape[mode], p)) if transp else g.random((p, shape[mode]))
            Y = S.ttm(V, mode, transp=transp)
            ref = orc.ttm_dense(A, V, mode, transp=transp)
            assert np.asarray(Y).shape == ref.shape
            assert relerr(np.asarray(Y), ref) < 1e-12, (mode, transp)
    # a chain over all modes but one (what tucker.hooi issues), continued on the dense kernels
    U = [g.random((I, 3)) for I in shape]
    Yc = S.ttm(U, 1, transp=True, without=True)
    assert relerr(np.asarray(Yc), orc.ttm_chain(A, U, 1, transp=True)) < 1e-12


def test_sparse_ttm_reference_known_answer(K, golden):
    """The reference's own sparse TTM test (test_sptensor.py:96-100), exact integers."""
    from sktensor import sptensor
    T = golden['kat1_T']
    Y = np.zeros((2, 4, 2))
    Y[:, :, 0] = [[22, 49, 76, 103], [28, 64, 100, 136]]
    Y[:, :, 1] = [[130, 157, 184, 211], [172, 208, 244, 280]]
    S = sptensor(T.nonzero(), T.flatten(), T.shape)
    Y2 = S.ttm(np.array([[1, 3, 5], [2, 4, 6]]), 0)
    assert (2, 4, 2) == Y2.shape and (Y == np.asarray(Y2)).all()


def test_sparse_hooi_matches_dense_oracle(K):
    from sktensor import tucker_hooi
    from sktensor.sptensor import fromarray
    rng = np.random.RandomState(8)
    shape, rank = (24, 20, 28), [4, 3, 5]
    G = rng.rand(*rank)
    X = G
    for m in range(3):
        X = np.moveaxis(np.tensordot(np.linalg.qr(rng.randn(shape[m], rank[m]))[0], X, axes=(1, m)), 0, m)
    X = (X + 0.05 * rng.randn(*shape)) * (rng.rand(*shape) > 0.5)
    np.random.seed(2)
    core, U = tucker_hooi(fromarray(X), rank, init='random', maxIter=8, conv=0)
    np.random.seed(2)
    core_o, U_o, fit_o, _ = orc.hooi(X, rank, init='random', maxIter=8, conv=0)
    nx = np.linalg.norm(X)
    fit = 1 - np.sqrt(max(nx ** 2 - np.linalg.norm(core) ** 2, 0.0)) / nx
    assert abs(fit - float(fit_o)) < 1e-8
    for m in range(3):
        assert relerr(U[m].dot(U[m].T), U_o[m].dot(U_o[m].T)) < 1e-7


@pytest.mark.parametrize('path', ['gram', 'lobpcg'])
def test_sparse_nvecs_matches_dense(K, golden, path, monkeypatch):
    """core.nvecs on an sptensor (core.py:280-283: eigsh of X_(n) X_(n)^T) without SciPy: small modes through the dense
    Gram + cluster eigensolver, long modes through the block Krylov iteration -- both forced here on the same tensor."""
    from sktensor.core import nvecs
    from sktensor.sptensor import fromarray
    monkeypatch.setenv('SKT_SPARSE_NVECS', path)
    rng = np.random.RandomState(1)
    shape, r = (60, 45, 50), 4
    F = [rng.rand(s, r) for s in shape]
    X = np.einsum('ir,jr,kr->ijk', *F) * (rng.rand(*shape) > 0.7)
    S = fromarray(X)
    for n in range(3):
        got = nvecs(S, n, 3)
        ref = orc.nvecs_dense(X, n, 3)
        assert got.shape == ref.shape
        assert relerr(got, ref) < 1e-7, (path, n)


def test_cp_als_sparse_default_init_golden(golden):
    """KAT-3 with the DEFAULT init ('nvecs'): the whole call runs without SciPy's eigsh."""
    from sktensor import cp_als
    from sktensor.sptensor import fromarray
    X1 = golden['cfg1_X']
    P, fit, itr, _ = cp_als(fromarray(X1 * (X1 > 0.7)), 3)
    assert itr == golden.j['cp_kat3_nvecs']['itr']
    assert abs(float(fit) - golden.j['cp_kat3_nvecs']['fit']) < 1e-6

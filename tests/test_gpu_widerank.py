"""Ranks beyond the single-CTA solve kernels (> 64): the reference's cp.als / ktensor.norm have no rank limit
(cp.py:46, ktensor.py:113-126).  The generic entry points switch to the wide path (csrc/cpals_wide.cu): Gram on the
FP64 tensor cores, pseudo-inverse through the cluster eigensolver, solve through the tiled GEMM."""
import numpy as np
import pytest
import scipy.linalg as sl

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


@pytest.mark.parametrize('rows,R', [(300, 65), (1000, 96), (5000, 128), (700, 200), (64, 70)])
def test_wide_gram_pinv_solve_normalise(K, dev, rows, R):
    rng = np.random.RandomState(rows + R)
    U = rng.rand(rows, R) - 0.3
    G = dev.to_host(K.gram(dev.to_device(U)))
    assert relerr(G, U.T.dot(U)) < 1e-13
    # Hadamard of three Grams (one of them rank deficient when rows < R) -> pinv with scipy's cut-off
    Gs = np.stack([(lambda A: A.T.dot(A))(rng.rand(max(rows // 4, 8), R)) for _ in range(3)])
    P, info = K.hadamard_pinv(dev.to_device(Gs), 1)
    Y = Gs[0] * Gs[2]
    ref = sl.pinv(Y)
    assert int(dev.to_host(info)[0]) == np.linalg.matrix_rank(Y, tol=R * np.finfo(float).eps * np.linalg.norm(Y, 2))
    assert relerr(dev.to_host(P), ref) < 1e-6 * max(1.0, np.linalg.cond(Y) * 1e-10)
    M = rng.rand(rows, R)
    Pm = 0.5 * (ref + ref.T)
    for first in (True, False):
        Unew, colstat = K.solve_stats(dev.to_device(M), dev.to_device(Pm), first)
        refU = M.dot(Pm)
        assert relerr(dev.to_host(Unew), refU) < 1e-12
        refstat = (refU ** 2).sum(axis=0) if first else refU.max(axis=0)
        assert relerr(dev.to_host(colstat), refstat) < 1e-12
        lam, Gn, ip = K.normalise_gram(Unew, colstat, first, Mip=dev.to_device(M))
        l = np.sqrt(refstat) if first else np.where(refstat < 1, 1.0, refstat)
        Un = refU / l
        assert relerr(dev.to_host(lam), l) < 1e-12
        assert relerr(dev.to_host(Unew), Un) < 1e-12
        assert relerr(dev.to_host(Gn), Un.T.dot(Un)) < 1e-12
        assert abs(float(dev.to_host(ip)[0]) - float((l * (Un * M).sum(axis=0)).sum())) < 1e-10 * abs(float((l * (Un * M).sum(axis=0)).sum()))


@pytest.mark.parametrize('shape,R,sparse', [((30, 28, 26), 80, False), ((24, 20, 18, 16), 66, False), ((40, 36, 30), 72, True)])
def test_cp_als_rank_above_64_matches_oracle(shape, R, sparse):
    from sktensor import cp_als, dtensor
    from sktensor.sptensor import fromarray
    rng = np.random.RandomState(R)
    X = rng.rand(*shape)
    if sparse:
        X = X * (X > 0.6)
    np.random.seed(3)
    P, fit, itr, _ = cp_als(fromarray(X) if sparse else dtensor(X), R, init='random', max_iter=4, conv=0)
    np.random.seed(3)
    if sparse:
        S = fromarray(X)
        Uo, lo, fo, io, _ = orc.cp_als((tuple(np.asarray(s) for s in S.subs), np.asarray(S.vals), S.shape), R, init='random',
                                       max_iter=4, conv=0)
    else:
        Uo, lo, fo, io, _ = orc.cp_als(X, R, init='random', max_iter=4, conv=0)
    assert itr == io
    assert abs(float(fit) - float(np.ravel(fo)[0])) < 1e-6


def test_ktensor_norm_many_components(K, dev):
    from sktensor import ktensor
    rng = np.random.RandomState(2)
    U = [rng.rand(s, 70) for s in (9, 8, 7)]
    lam = rng.rand(70)
    assert abs(ktensor(U, lam).norm() - orc.kruskal_norm(U, lam)) < 1e-10 * orc.kruskal_norm(U, lam)


def test_rank_limit_is_a_clear_error():
    from sktensor import cp_als, dtensor
    with pytest.raises(ValueError):
        cp_als(dtensor(np.random.rand(4, 4, 4)), 600, init='random', max_iter=1)

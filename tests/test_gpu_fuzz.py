"""Randomised end-to-end parity: the public cp_als / tucker_hooi calls against the CPU oracle over many small,
awkward configurations (odd sizes -> cp.async kernel, every tensor order 2..5, rank 1..64, dense and sparse,
every update path: cluster, fused, exchange).  Tolerances: iteration count exact, fit within 1e-8."""
import numpy as np
import pytest

from conftest import relerr
from oracle import sktensor_oracle as orc

pytestmark = pytest.mark.gpu


def _configs():
    rng = np.random.RandomState(2024)
    out = []
    for k in range(28):
        N = int(rng.choice([2, 3, 3, 3, 4, 4, 5]))
        hi = {2: 90, 3: 40, 4: 16, 5: 9}[N]
        shape = tuple(int(rng.randint(1 if k % 7 == 0 else 2, hi)) for _ in range(N))
        R = int(rng.choice([1, 2, 3, 5, 8, 16, 17, 32, 33, 64]))
        sparse = bool(k % 3 == 2)
        out.append((k, shape, R, sparse))
    return out


@pytest.mark.parametrize('k,shape,R,sparse', _configs())
def test_cp_als_random_config_matches_oracle(k, shape, R, sparse):
    from sktensor import cp_als, dtensor
    from sktensor.sptensor import fromarray
    rng = np.random.RandomState(1000 + k)
    X = rng.rand(*shape)
    if sparse:
        X = X * (rng.rand(*shape) > 0.6)
        if not X.any():
            X.flat[0] = 1.0
    T = fromarray(X) if sparse else dtensor(X)
    np.random.seed(k)
    P, fit, itr, _ = cp_als(T, R, init='random', max_iter=12, conv=1e-9)
    np.random.seed(k)
    Xo = (tuple(np.asarray(s_) for s_ in T.subs), np.asarray(T.vals, dtype=float), tuple(T.shape)) if sparse else X
    Uo, lo, fo, io, _ = orc.cp_als(Xo, R, init='random', max_iter=12, conv=1e-9)
    fo = float(np.ravel(fo)[0])
    if not np.isfinite(fo):          # degenerate draw (zero column on iteration 0): NaN on both sides
        assert not np.isfinite(fit)
        return
    assert itr == io, (shape, R, sparse, itr, io)
    assert abs(fit - fo) < 1e-8 * max(1.0, abs(fo)), (shape, R, sparse, fit, fo)
    for m in range(len(shape)):
        assert P.U[m].shape == (shape[m], R)


@pytest.mark.parametrize('k', range(8))
def test_hooi_random_config_matches_oracle(k):
    from sktensor import dtensor, tucker_hooi
    rng = np.random.RandomState(500 + k)
    N = int(rng.choice([3, 3, 4]))
    shape = tuple(int(rng.randint(4, 30 if N == 3 else 14)) for _ in range(N))
    rank = [int(rng.randint(1, max(2, s // 2 + 1))) for s in shape]
    # a mode rank above the product of the other ranks makes that mode's Gram rank deficient: the surplus eigenvectors
    # span a null space in which LAPACK's (and any solver's) basis is arbitrary, and the next sweep depends on it
    rank = [min(r, int(np.prod([rank[j] for j in range(N) if j != m]))) for m, r in enumerate(rank)]
    X = rng.rand(*shape)
    np.random.seed(k)
    core, U = tucker_hooi(dtensor(X), rank, init='random', maxIter=6, conv=0)
    np.random.seed(k)
    core_o, U_o = orc.hooi(X, rank, init='random', maxIter=6, conv=0)[:2]
    assert core.shape == tuple(rank)
    # unconverged HOOI on a random tensor has nearly degenerate trailing eigenvalues, so rounding differences are
    # amplified from sweep to sweep: compare the fit at north_star's tolerance (1e-6), not the raw core
    nx = np.linalg.norm(X)
    fit = 1 - np.sqrt(max(nx ** 2 - np.linalg.norm(core) ** 2, 0.0)) / nx
    fit_o = 1 - np.sqrt(max(nx ** 2 - np.linalg.norm(core_o) ** 2, 0.0)) / nx
    assert abs(fit - fit_o) < 1e-6, (shape, rank, fit, fit_o)
    for m in range(N):
        assert relerr(U[m].T.dot(U[m]), np.eye(rank[m])) < 1e-10

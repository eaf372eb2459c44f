"""Opt-in sliced contraction on the INT8 tensor cores (csrc/slice_gemm.cu, `skt_planes_*`): the two partial contractions
of the 3-way dimension tree, FP64-accurate by error-free slicing.  The kernel is checked against NumPy's FP64 product of
the same unfolding (the reference's khatrirao + dot on a 2-way view, sktensor/dtensor.py:117-137) with the MTTKRP parity
measure and tolerance of SURVEY 8(d): max|a - b| / max|b| < 1e-10."""
import numpy as np
import pytest

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


def _draw(rng, shape, dist):
    if dist == 'uniform':
        return rng.rand(*shape)
    if dist == 'normal':
        return rng.randn(*shape)
    if dist == 'wide':                                  # six decades of dynamic range, both signs
        return rng.randn(*shape) * 10.0 ** rng.uniform(-3, 3, size=shape)
    raise ValueError(dist)


def _ref(X, mode, U):
    if mode == 0:
        return X.reshape(X.shape[0], -1).T.dot(U)
    return X.reshape(-1, X.shape[-1]).dot(U)


SHAPES = [((16, 16, 16), 4), ((48, 20, 64), 7), ((130, 36, 208), 32), ((512, 8, 512), 32), ((96, 100, 48), 1),
          ((33, 16, 400), 20), ((400, 16), 32), ((8, 6, 10, 32), 5)]


@pytest.mark.parametrize('shape,R', SHAPES)
@pytest.mark.parametrize('dist', ['uniform', 'normal'])
@pytest.mark.parametrize('S', [6, 5])
def test_planes_ttm_matches_numpy(K, dev, shape, R, dist, S):
    rng = np.random.RandomState(sum(shape) + R + S)
    X = _draw(rng, shape, dist)
    Xd = dev.to_device(X)
    P = K.PlanesHandle(Xd, shape, S)
    try:
        assert P.device_bytes() >= S * X.size
        for mode in (0, len(shape) - 1):
            U = _draw(rng, (shape[mode], R), dist)
            if not K.PlanesHandle.supported(shape, mode, R):
                with pytest.raises(Exception):
                    P.ttm(mode, dev.to_device(U), R)
                continue
            T = dev.to_host(P.ttm(mode, dev.to_device(U), R))
            ref = _ref(X, mode, U)
            assert T.shape == ref.shape
            # bound of the header: 2^-(8S-1) (max|X| sum|U| + max|U| sum|X|) per entry, far below the parity tolerance
            bound = 2.0 ** -(8 * S - 1) * (np.abs(X).max() * np.abs(U).sum(axis=0).max() +
                                            np.abs(U).max() * np.abs(X).sum(axis=mode).max())
            assert np.abs(T - ref).max() <= 4 * bound + 1e-15 * np.abs(ref).max()
            assert relerr(T, ref) < (1e-12 if S == 6 else 1e-10)
    finally:
        P.close()


def test_planes_wide_dynamic_range_stays_within_the_bound(K, dev):
    rng = np.random.RandomState(5)
    shape = (64, 48, 128)
    X = _draw(rng, shape, 'wide')
    P = K.PlanesHandle(dev.to_device(X), shape, 6)
    try:
        for mode in (0, 2):
            U = _draw(rng, (shape[mode], 16), 'wide')
            T = dev.to_host(P.ttm(mode, dev.to_device(U), 16))
            ref = _ref(X, mode, U)
            bound = 2.0 ** -47 * (np.abs(X).max() * np.abs(U).sum(axis=0) + np.abs(U).max(axis=0) * np.abs(X).sum(axis=mode).max())
            assert (np.abs(T - ref) <= 4 * bound[None, :] + 1e-15 * np.abs(ref).max()).all()
    finally:
        P.close()


def test_planes_exact_on_integers_and_zero(K, dev):
    rng = np.random.RandomState(11)
    shape = (32, 16, 64)
    X = rng.randint(-1000, 1000, size=shape).astype(float)
    X[3] = 0.0
    P = K.PlanesHandle(dev.to_device(X), shape, 6)
    Z = K.PlanesHandle(dev.to_device(np.zeros(shape)), shape, 6)
    try:
        for mode in (0, 2):
            U = rng.randint(-50, 50, size=(shape[mode], 8)).astype(float)
            U[:, 2] = 0.0                                # a zero column keeps scale 1 and gives exact zeros
            T = dev.to_host(P.ttm(mode, dev.to_device(U), 8))
            assert np.array_equal(T, _ref(X, mode, U))   # small integers: every digit product is exact
            assert not dev.to_host(Z.ttm(mode, dev.to_device(U), 8)).any()
    finally:
        P.close()
        Z.close()


def test_planes_non_finite_input_poisons_the_output(K, dev):
    shape = (16, 16, 32)
    X = np.random.RandomState(2).rand(*shape)
    U = np.random.RandomState(3).rand(32, 4)
    Xbad = X.copy()
    Xbad[1, 2, 3] = np.inf
    P = K.PlanesHandle(dev.to_device(Xbad), shape, 6)
    try:
        assert np.isnan(dev.to_host(P.ttm(2, dev.to_device(U), 4))).all()
    finally:
        P.close()
    P = K.PlanesHandle(dev.to_device(X), shape, 6)
    try:
        Ubad = U.copy()
        Ubad[5, 1] = np.nan
        T = dev.to_host(P.ttm(2, dev.to_device(Ubad), 4))
        assert np.isnan(T[:, 1]).all() and np.isfinite(T[:, [0, 2, 3]]).all()
    finally:
        P.close()


def test_planes_rejects_what_it_does_not_support(K, dev):
    shape = (24, 16, 32)
    Xd = dev.to_device(np.ones(shape))
    with pytest.raises(Exception):
        K.PlanesHandle(Xd, shape, 4)
    P = K.PlanesHandle(Xd, shape, 6)
    try:
        assert not K.PlanesHandle.supported(shape, 1, 8)          # a middle mode
        assert not K.PlanesHandle.supported(shape, 2, 33)         # rank beyond one slice block
        assert not K.PlanesHandle.supported((24, 16, 600), 2, 8)  # contraction longer than the resident factor planes
        assert not K.PlanesHandle.supported((24, 16, 30), 2, 8)   # rows of the planes not 16-byte aligned
        with pytest.raises(Exception):
            P.ttm(1, dev.to_device(np.ones((16, 8))), 8)
    finally:
        P.close()


# ---------------------------------------------------------------------------------------------------------------------
# cp_als with the sliced passes switched on (SKT_DENSE_I8): same iterates as the FP64 engine and as the oracle
@pytest.fixture
def sliced(monkeypatch):
    monkeypatch.setenv('SKT_DENSE_I8', '6')


def _cp(X, R, U0, **kw):
    from sktensor import cp_als, dtensor
    return cp_als(dtensor(X), R, init=[u.copy() for u in U0], **kw)


@pytest.mark.parametrize('shape,R', [((64, 48, 32), 8), ((32, 80, 64), 32), ((208, 16, 16), 5), ((16, 16, 304), 12)])
def test_cp_als_sliced_matches_fp64_engine_and_oracle(monkeypatch, shape, R):
    from oracle import sktensor_oracle as orc
    rng = np.random.RandomState(sum(shape))
    X = rng.rand(*shape)
    U0 = [rng.rand(d, R) for d in shape]
    P0, fit0, itr0, _ = _cp(X, R, U0, max_iter=12, conv=0)
    monkeypatch.setenv('SKT_DENSE_I8', '6')
    P1, fit1, itr1, _ = _cp(X, R, U0, max_iter=12, conv=0)
    assert itr1 == itr0
    assert abs(fit1 - fit0) < 1e-9
    for a, b in zip(P1.U, P0.U):
        assert relerr(a, b) < 1e-8
    assert relerr(P1.lmbda, P0.lmbda) < 1e-8
    Uo, lamo, fito, itro = orc.cp_als(X, R, init=[u.copy() for u in U0], max_iter=12, conv=0)[:4]
    assert abs(fit1 - fito) < 1e-6


def test_cp_als_sliced_engine_really_takes_the_sliced_passes(sliced):
    from sktensor import _kernels as K, _parallel as par, dtensor
    rng = np.random.RandomState(1)
    X = rng.rand(32, 32, 32)
    U = [rng.rand(32, 4) for _ in range(3)]
    eng = par.CpAlsEngine(dtensor(X), 4, U)
    try:
        assert eng.i8 is not None and eng.i8.nslices == 6
        eng.mttkrp(0)
    finally:
        eng.release()
    assert eng.i8 is None


def test_cp_als_sliced_falls_back_to_fp64_when_the_shape_is_not_supported(sliced):
    rng = np.random.RandomState(2)
    X = rng.rand(20, 18, 30)                               # plane rows not 16-byte aligned
    U0 = [rng.rand(d, 3) for d in X.shape]
    P, fit, itr, _ = _cp(X, 3, U0, max_iter=5, conv=0)
    assert np.isfinite(fit)


# ---------------------------------------------------------------------------------------------------------------------
# the general tree pass (skt_planes_pass): X as [P x Q], one long GEMM against the Khatri-Rao product of a group's factors
def _kr(mats):
    out = mats[0]
    for M in mats[1:]:
        out = (out[:, None, :] * M[None, :, :]).reshape(-1, M.shape[1])
    return out


PASS_CASES = [((16, 16, 16), 1, 4), ((16, 16, 16), 2, 4), ((12, 20, 8, 16), 2, 64), ((40, 6, 10, 32), 2, 33),
              ((160, 20, 160), 2, 64), ((9, 64, 80), 1, 17), ((24, 40, 608), 2, 32), ((6, 5, 4, 8, 16), 3, 48),
              ((2100, 16), 1, 8), ((130, 20, 34, 32), 2, 20)]


@pytest.mark.parametrize('shape,split,R', PASS_CASES)
@pytest.mark.parametrize('S', [6, 5])
def test_planes_tree_pass_matches_numpy(K, dev, shape, split, R, S):
    rng = np.random.RandomState(sum(shape) + R + S + split)
    X = rng.rand(*shape) - 0.25
    U = [rng.rand(d, R) - 0.4 for d in shape]
    Ud = [dev.to_device(u) for u in U]
    assert K.PlanesHandle.pass_supported(shape, split, R)
    P = K.PlanesHandle(dev.to_device(X), shape, S)
    try:
        Pn = int(np.prod(shape[:split]))
        X2 = X.reshape(Pn, -1)
        TA = dev.to_host(P.tree_pass(split, 0, Ud, R))
        refA = X2.dot(_kr(U[split:]))
        TB = dev.to_host(P.tree_pass(split, 1, Ud, R))
        refB = X2.T.dot(_kr(U[:split]))
        tol = 1e-12 if S == 6 else 1e-10
        assert TA.shape == refA.shape and TB.shape == refB.shape
        assert relerr(TA, refA) < tol
        assert relerr(TB, refB) < tol
    finally:
        P.close()


def test_planes_tree_pass_long_contraction_crosses_the_int32_flush(K, dev):
    # 20480 products per output: two K segments at least (an INT32 accumulator holds 16384 of them safely); worst-case digits
    shape, R = (20480, 16), 8
    X = np.full(shape, -1.0)
    X[::3] = 0.999
    U = [np.full((shape[0], R), -1.0), np.full((shape[1], R), 1.0)]
    P = K.PlanesHandle(dev.to_device(X), shape, 6)
    try:
        T = dev.to_host(P.tree_pass(1, 1, [dev.to_device(u) for u in U], R))
        assert relerr(T, X.T.dot(U[0])) < 1e-12
    finally:
        P.close()


@pytest.mark.parametrize('shape,R', [((12, 16, 10, 16), 6), ((20, 12, 16, 16), 40), ((48, 16, 640), 16), ((64, 32, 32), 50)])
def test_cp_als_sliced_general_passes_match_fp64_engine(monkeypatch, shape, R):
    rng = np.random.RandomState(sum(shape) + R)
    X = rng.rand(*shape)
    U0 = [rng.rand(d, R) for d in shape]
    P0, fit0, itr0, _ = _cp(X, R, U0, max_iter=8, conv=0)
    monkeypatch.setenv('SKT_DENSE_I8', '6')
    from sktensor import _parallel as par, dtensor
    eng = par.CpAlsEngine(dtensor(X), R, [u.copy() for u in U0])
    try:
        assert eng.i8 is not None and eng.i8_general
    finally:
        eng.release()
    P1, fit1, itr1, _ = _cp(X, R, U0, max_iter=8, conv=0)
    assert abs(fit1 - fit0) < 1e-9
    for a, b in zip(P1.U, P0.U):
        assert relerr(a, b) < 1e-7


# ---------------------------------------------------------------------------------------------------------------------
# the second stage folded into the epilogue (skt_planes_ttm_fused): slabs whose sum is the finished MTTKRP
@pytest.mark.parametrize('shape,R', [((16, 128, 128), 4), ((40, 256, 384), 32), ((130, 128, 256), 7), ((512, 512, 128), 32)])
@pytest.mark.parametrize('S', [6, 5])
def test_planes_ttm_fused_slabs_sum_to_the_mttkrp(K, dev, shape, R, S):
    rng = np.random.RandomState(sum(shape) + R + S)
    X = rng.rand(*shape) - 0.3
    U = [rng.rand(d, R) - 0.2 for d in shape]
    Ud = [dev.to_device(u) for u in U]
    P = K.PlanesHandle(dev.to_device(X), shape, S)
    tol = 1e-12 if S == 6 else 1e-10
    try:
        assert P.fused_slabs(2, R) == (0, 0)                     # only the first-mode contraction has a fused form
        # first-mode contraction: only slabs of the mode-2 MTTKRP
        ns, rows = P.fused_slabs(0, R)
        assert ns >= 1 and rows == shape[2]
        slabs = dev.zeros((ns, rows, R)) + 7.0
        P.ttm_fused(0, Ud[0], Ud[1], R, slabs)
        M2 = np.einsum('ijk,ir,jr->kr', X, U[0], U[1])
        assert relerr(dev.to_host(slabs).sum(axis=0), M2) < tol
    finally:
        P.close()


def test_planes_fused_needs_multiples_of_128(K, dev):
    shape = (32, 96, 160)
    P = K.PlanesHandle(dev.to_device(np.ones(shape)), shape, 6)
    try:
        assert P.fused_slabs(2, 8) == (0, 0) and P.fused_slabs(0, 8) == (0, 0)
    finally:
        P.close()


@pytest.mark.parametrize('shape,R', [((48, 128, 128), 8), ((128, 256, 128), 32)])
def test_cp_als_sliced_fused_matches_unfused_and_fp64(monkeypatch, shape, R):
    from sktensor import _parallel as par, dtensor
    rng = np.random.RandomState(sum(shape))
    X = rng.rand(*shape)
    U0 = [rng.rand(d, R) for d in shape]
    P0, fit0, itr0, _ = _cp(X, R, U0, max_iter=10, conv=0)
    monkeypatch.setenv('SKT_DENSE_I8', '6')
    eng = par.CpAlsEngine(dtensor(X), R, [u.copy() for u in U0])
    try:
        assert eng.i8 is not None and eng.i8_fused[2] and not eng.i8_fused[0]
    finally:
        eng.release()
    P1, fit1, itr1, _ = _cp(X, R, U0, max_iter=10, conv=0)
    monkeypatch.setenv('SKT_I8_NO_FUSED', '1')
    P2, fit2, itr2, _ = _cp(X, R, U0, max_iter=10, conv=0)
    assert abs(fit1 - fit0) < 1e-9 and abs(fit2 - fit0) < 1e-9
    for a, b, c in zip(P1.U, P0.U, P2.U):
        assert relerr(a, b) < 1e-8 and relerr(c, b) < 1e-8


@pytest.mark.parametrize('top', [1.0, 2.0 ** -30, 0.9921875, 0.9921875 * (1 + 2.0 ** -40), 0.9921875 * (1 - 2.0 ** -40), 1 - 2.0 ** -50,
                                 3.0, 1e-300, 1e300])
@pytest.mark.parametrize('S', [6, 5])
def test_planes_scale_boundaries(K, dev, top, S):
    """The largest magnitude decides the power-of-two scale; values at the edge of the balanced-digit range (127.x * 256^(S-1))
    must neither overflow the top digit nor lose more than the stated bound."""
    rng = np.random.RandomState(3)
    shape = (16, 16, 32)
    X = (rng.rand(*shape) - 0.5) * top
    X[0, 0, 0], X[1, 2, 3], X[5, 5, 5] = top, -top, top * (1 - 2.0 ** -20)
    U = rng.rand(32, 4) - 0.5
    U[3, 1], U[4, 2] = -1.0, 0.9921875
    P = K.PlanesHandle(dev.to_device(X), shape, S)
    try:
        T = dev.to_host(P.ttm(2, dev.to_device(U), 4))
        ref = X.reshape(-1, 32).dot(U)
        bound = 2.0 ** -(8 * S - 1) * (np.abs(X).max() * np.abs(U).sum(axis=0).max() + np.abs(U).max() * np.abs(X).sum(axis=2).max())
        assert np.isfinite(T).all()
        assert np.abs(T - ref).max() <= 4 * bound + 4e-16 * np.abs(ref).max()
    finally:
        P.close()


def test_cp_als_sliced_with_five_slices(monkeypatch):
    """SKT_DENSE_I8=5 (40-bit operands): still far inside the reference tolerances (MTTKRP 1e-10, fit 1e-6)."""
    from oracle import sktensor_oracle as orc
    rng = np.random.RandomState(9)
    shape, R = (64, 128, 128), 16
    X = rng.rand(*shape)
    U0 = [rng.rand(d, R) for d in shape]
    monkeypatch.setenv('SKT_DENSE_I8', '5')
    P1, fit1, itr1, _ = _cp(X, R, U0, max_iter=10, conv=0)
    Uo, lamo, fito, itro = orc.cp_als(X, R, init=[u.copy() for u in U0], max_iter=10, conv=0)[:4]
    assert itr1 == itro and abs(fit1 - fito) < 1e-8
    for a, b in zip(P1.U, Uo):
        assert relerr(a, b) < 1e-6


def test_planes_fuzz_random_shapes_splits_and_ranks(K, dev):
    """Seeded sweep over the corner structure of the general pass and the resident kernel: ragged row tiles, contraction tails
    that are not a multiple of 32 or 128 bytes, 1 .. 64 columns (partial 32- and 64-column blocks), split-K with few rows,
    2- to 5-way tensors, both groups, both slice counts."""
    rng = np.random.RandomState(20261020)
    tail_dims = [16, 32, 48, 80, 144, 160, 272, 400, 528]
    done = 0
    for trial in range(48):
        ndim = int(rng.randint(2, 6))
        split = int(rng.randint(1, ndim))
        shape = [int(rng.choice([1, 2, 3, 5, 7, 9, 12, 17, 33])) for _ in range(ndim)]
        shape[-1] = int(rng.choice(tail_dims))                         # Q = prod(shape[split:]) becomes a multiple of 16
        if rng.rand() < 0.4:
            shape[0] = int(rng.choice([130, 257, 600, 1500]))          # many row tiles on one side ...
        if rng.rand() < 0.3 and split < ndim - 1:
            shape[split] = int(rng.choice([64, 129, 300]))             # ... or a long contraction for group 0
        if np.prod(shape) > 3e6:
            shape[0] = max(1, int(3e6 // np.prod(shape[1:])))
        shape = tuple(shape)
        R = int(rng.choice([1, 2, 7, 16, 31, 32, 33, 40, 63, 64]))
        S = int(rng.choice([5, 6]))
        if not K.PlanesHandle.pass_supported(shape, split, R):
            continue
        X = rng.randn(*shape) if rng.rand() < 0.5 else rng.rand(*shape) - 0.3
        U = [rng.rand(d, R) - 0.4 for d in shape]
        Ud = [dev.to_device(u) for u in U]
        P = K.PlanesHandle(dev.to_device(X), shape, S)
        try:
            X2 = X.reshape(int(np.prod(shape[:split])), -1)
            tol = 1e-12 if S == 6 else 1e-10
            TA = dev.to_host(P.tree_pass(split, 0, Ud, R))
            TB = dev.to_host(P.tree_pass(split, 1, Ud, R))
            refA, refB = X2.dot(_kr(U[split:])), X2.T.dot(_kr(U[:split]))
            assert relerr(TA, refA) < tol, (trial, shape, split, R, S, 'group 0')
            assert relerr(TB, refB) < tol, (trial, shape, split, R, S, 'group 1')
            for mode in (0, ndim - 1):
                if K.PlanesHandle.supported(shape, mode, R):
                    T = dev.to_host(P.ttm(mode, Ud[mode], R))
                    ref = X.reshape(shape[0], -1).T.dot(U[0]) if mode == 0 else X.reshape(-1, shape[-1]).dot(U[-1])
                    assert relerr(T, ref) < tol, (trial, shape, mode, R, S, 'ttm')
            done += 1
        finally:
            P.close()
    assert done >= 30

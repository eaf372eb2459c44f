"""GPU parity tests for the MTTKRP kernels, through the C ABI (ctypes) and through the
drop-in methods.  Tolerance: north_star's 1e-10 relative (max|M - M_ref| / max|M_ref|);
integer-valued cases must be exact."""
import numpy as np
import pytest

from conftest import relerr
from oracle import sktensor_oracle as orc

pytestmark = pytest.mark.gpu

TOL = 1e-10


@pytest.fixture(scope='module')
def K():
    from sktensor import _device as dev, _kernels as K
    dev.require_cuda()
    return K


@pytest.fixture(scope='module')
def dev():
    from sktensor import _device as dev
    return dev


@pytest.fixture(params=['tma', 'cpasync'])
def dense_path(request):
    """Run a test on both dense kernels: TMA-staged (default) and the cp.async one used for
    tensors whose rows are not 16-byte aligned (forced here through SKT_DENSE_NO_TMA)."""
    import os
    os.environ['SKT_DENSE_NO_TMA'] = '1' if request.param == 'cpasync' else '0'
    yield request.param
    os.environ['SKT_DENSE_NO_TMA'] = '0'


@pytest.fixture(params=['staged', 'registers'])
def sparse_path(request):
    """Both sparse kernels: cp.async-staged gathers (default for even ranks) and the register-gather variant."""
    import os
    os.environ['SKT_COO_NO_STAGED'] = '1' if request.param == 'registers' else '0'
    yield request.param
    os.environ['SKT_COO_NO_STAGED'] = '0'


def gpu_mttkrp_dense(K, dev, X, U, n, naive=False):
    R = next(u for m, u in enumerate(U) if m != n and u is not None).shape[1]
    Ud = [None if (u is None or m == n) else dev.to_device(u) for m, u in enumerate(U)]
    M = K.mttkrp_dense(dev.to_device(X), X.shape, Ud, R, n, naive=naive)
    return dev.to_host(M)


def test_native_library_is_loaded(K):
    from sktensor import _lib
    maps = open('/proc/self/maps').read()
    assert 'libsktensor_b200.so' in maps
    sm, major, minor, smem = [__import__('ctypes').c_int() for _ in range(3)] + [__import__('ctypes').c_size_t()]
    import ctypes
    _lib.check(_lib.load().skt_device_info(ctypes.byref(sm), ctypes.byref(major), ctypes.byref(minor), ctypes.byref(smem)))
    assert major.value == 10 and sm.value >= 100 and smem.value >= 227 * 1024


def test_kat1_exact_integers(K, dev, golden, dense_path):
    T = golden['kat1_T']
    U = golden.factors('kat1', 3)
    for n in range(3):
        assert (gpu_mttkrp_dense(K, dev, T, U, n) == golden['kat1_M%d' % n]).all()
        assert (gpu_mttkrp_dense(K, dev, T, U, n, naive=True) == golden['kat1_M%d' % n]).all()


def test_dense_golden_all_modes(K, dev, golden, dense_path):
    for name, c in golden.j['dense_cases'].items():
        N = len(c['shape'])
        X = golden['mtd_%s_X' % name]
        U = golden.factors('mtd_%s' % name, N)
        for n in range(N):
            Uin = list(U)
            Uin[n] = None                                   # U[mode] may be None (cp.py:194,143)
            ref = golden['mtd_%s_M%d' % (name, n)]
            assert relerr(gpu_mttkrp_dense(K, dev, X, Uin, n), ref) < TOL, (name, n)
            assert relerr(gpu_mttkrp_dense(K, dev, X, Uin, n, naive=True), ref) < TOL, (name, n)


@pytest.mark.parametrize('shape,R', [
    ((7, 5, 3), 1), ((8, 9, 10), 3), ((31, 17, 29), 8), ((64, 64, 64), 16), ((40, 33, 48), 31),
    ((129, 20, 34), 33), ((20, 130, 66), 64), ((24, 18, 20), 65), ((12, 14, 16), 100),
    ((6, 7, 8, 9), 5), ((16, 12, 10, 14), 32), ((5, 4, 6, 3, 7), 4), ((300, 2), 6), ((2, 300), 6),
    ((1, 50, 40), 4), ((50, 1, 40), 4), ((50, 40, 1), 4), ((3, 3, 3, 3, 3, 3), 2),
    ((130, 70, 36), 32), ((36, 130, 70), 16), ((18, 22, 14, 20), 64), ((160, 12, 6, 40), 24),
])
def test_dense_random_shapes_vs_oracle(K, dev, shape, R, dense_path):
    rng = np.random.RandomState(hash((shape, R)) % (2 ** 31))
    X = rng.rand(*shape) - 0.3
    U = [rng.rand(s, R) - 0.5 for s in shape]
    for n in range(len(shape)):
        ref = orc.mttkrp_dense(X, U, n)
        got = gpu_mttkrp_dense(K, dev, X, U, n)
        assert got.shape == ref.shape
        assert relerr(got, ref) < TOL, (shape, R, n)


def test_dense_layout_and_dtype_independence(K, dev):
    # F-order, strided and float32 / integer inputs give the same answer as the reference (App. A.15/16)
    from sktensor import dtensor
    rng = np.random.RandomState(3)
    X = rng.rand(9, 10, 11)
    U = [rng.rand(s, 4) for s in X.shape]
    ref = [orc.mttkrp_dense(X, U, n) for n in range(3)]
    big = rng.rand(18, 10, 11)
    big[::2] = X
    for Xv in (np.asfortranarray(X), big[::2]):
        for n in range(3):
            got = dtensor(Xv).uttkrp(U, n)
            assert relerr(got, ref[n]) < TOL
    Xi = rng.randint(-5, 6, size=(6, 7, 8))
    Ui = [rng.randint(-3, 4, size=(s, 3)).astype(float) for s in Xi.shape]
    for n in range(3):
        assert (np.asarray(dtensor(Xi).uttkrp(Ui, n)) == orc.mttkrp_dense(Xi.astype(float), Ui, n)).all()
    X32 = X.astype(np.float32)
    assert relerr(dtensor(X32).uttkrp(U, 1), orc.mttkrp_dense(X32.astype(float), U, 1)) < TOL


def test_dense_method_contract(K, dev, golden):
    from sktensor import dtensor, unfolded_dtensor
    X = golden['cfg1_X']
    rng = np.random.RandomState(0)
    U = [rng.rand(s, 3) for s in X.shape]
    keep = [u.copy() for u in U]
    Xk = X.copy()
    M = dtensor(X).uttkrp([None] + U[1:], 0)
    assert isinstance(M, (unfolded_dtensor, np.ndarray)) and M.shape == (10, 3) and M.dtype == np.float64
    assert (X == Xk).all() and all((a == b).all() for a, b in zip(U, keep))     # inputs not mutated
    assert relerr(M, orc.mttkrp_dense(X, U, 0)) < TOL
    with pytest.raises(ValueError):
        dtensor(X).uttkrp([U[0], U[1][:, :2], U[2]], 0)
    with pytest.raises(ValueError):
        dtensor(X).uttkrp(U, 3)
    # cached device copy gives the same result
    Xc = dtensor(X).cache_on_device()
    assert (np.asarray(Xc.uttkrp(U, 2)) == np.asarray(dtensor(X).uttkrp(U, 2))).all()


def test_dense_host_buffer_abi(K, golden):
    X = golden['mtd_d3b_X']
    U = golden.factors('mtd_d3b', 3)
    for n in range(3):
        assert relerr(K.mttkrp_dense_host(X, U, 32, n), golden['mtd_d3b_M%d' % n]) < TOL


def test_dense_deterministic_and_linear(K, dev, dense_path):
    """Size-independent properties at a size the oracle still checks in seconds (256^3, R=32)."""
    rng = np.random.RandomState(11)
    X = rng.rand(256, 256, 256)
    U = [rng.rand(256, 32) for _ in range(3)]
    Xd = dev.to_device(X)
    Ud = [dev.to_device(u) for u in U]
    for n in range(3):
        a = dev.to_host(K.mttkrp_dense(Xd, X.shape, Ud, 32, n))
        b = dev.to_host(K.mttkrp_dense(Xd, X.shape, Ud, 32, n))
        assert (a == b).all()                                     # bitwise reproducible
        assert relerr(a, orc.mttkrp_dense(X, U, n)) < TOL
        c = dev.to_host(K.mttkrp_dense(Xd, X.shape, Ud, 32, n, naive=True))
        assert relerr(a, c) < TOL
    # linearity in X and in one factor
    Y = rng.rand(*X.shape)
    Yd = dev.to_device(Y)
    Zd = dev.to_device(2.0 * X - 0.5 * Y)
    mx = dev.to_host(K.mttkrp_dense(Xd, X.shape, Ud, 32, 1))
    my = dev.to_host(K.mttkrp_dense(Yd, X.shape, Ud, 32, 1))
    mz = dev.to_host(K.mttkrp_dense(Zd, X.shape, Ud, 32, 1))
    assert relerr(mz, 2.0 * mx - 0.5 * my) < 1e-12


def test_dense_full_size_cfg2_properties(K, dev):
    """BASELINE config 2 (512^3, rank 32): cross-check against the independent naive kernel and
    the all-ones identity  sum_i M[i, r] * U_n[i, r] = <X, [[U]]>  shared by every mode."""
    rng = np.random.RandomState(0)
    X = rng.rand(512, 512, 512)
    U = [rng.rand(512, 32) for _ in range(3)]
    Xd = dev.to_device(X)
    del X
    Ud = [dev.to_device(u) for u in U]
    ips = []
    for n in range(3):
        M = K.mttkrp_dense(Xd, (512, 512, 512), Ud, 32, n)
        Mn = K.mttkrp_dense(Xd, (512, 512, 512), Ud, 32, n, naive=True)
        assert relerr(dev.to_host(M), dev.to_host(Mn)) < TOL
        ips.append(float(dev.to_host(K.weighted_dot(Ud[n], M))[0]))
    assert abs(ips[0] - ips[1]) / abs(ips[0]) < 1e-12 and abs(ips[0] - ips[2]) / abs(ips[0]) < 1e-12


# ------------------------------------------------------------------------------ sparse
def gpu_mttkrp_coo(K, dev, subs, vals, shape, U, n, atomic=False):
    coo = K.CooHandle(subs, vals, shape)
    R = next(u for m, u in enumerate(U) if m != n and u is not None).shape[1]
    Ud = [None if (u is None or m == n) else dev.to_device(u) for m, u in enumerate(U)]
    M = dev.to_host(K.mttkrp_coo(coo, Ud, R, n, atomic=atomic))
    coo.close()
    return M


def test_sparse_golden_all_modes(K, dev, golden, sparse_path):
    for name, c in golden.j['sp_cases'].items():
        N = len(c['shape'])
        subs = [golden['mts_%s_sub%d' % (name, m)] for m in range(N)]
        vals = golden['mts_%s_vals' % name]
        U = golden.factors('mts_%s' % name, N)
        for n in range(N):
            Uin = list(U)
            Uin[n] = None
            ref = golden['mts_%s_M%d' % (name, n)]
            got = gpu_mttkrp_coo(K, dev, subs, vals, c['shape'], Uin, n)
            assert got.dtype == np.float64 and got.shape == ref.shape
            assert relerr(got, ref) < TOL, (name, n)
            assert relerr(gpu_mttkrp_coo(K, dev, subs, vals, c['shape'], Uin, n, atomic=True), ref) < TOL
    assert (gpu_mttkrp_coo(K, dev, [golden['mts_dup_sub%d' % m] for m in range(3)], golden['mts_dup_vals'],
                           (3, 3, 2), golden.factors('mts_dup', 3), 0) == [[9, 24], [9, 24], [135, 216]]).all()


@pytest.mark.parametrize('shape,nnz,R,alpha', [
    ((100, 80, 60), 5000, 16, 3.0), ((1000, 50, 20), 30000, 3, 3.0), ((64, 64, 64), 70000, 32, 1.0),
    ((50, 40, 30, 20), 8000, 8, 2.0), ((20000, 300, 7), 60000, 64, 3.0), ((9, 8, 7, 6, 5), 3000, 5, 1.0),
    ((300, 200), 4000, 4, 2.0), ((40, 30, 20), 1, 16, 1.0), ((5, 200000, 3), 50000, 33, 3.0),
])
def test_sparse_random_vs_oracle(K, dev, shape, nnz, R, alpha, sparse_path):
    g = np.random.default_rng(abs(hash((shape, nnz, R))) % (2 ** 31))
    subs = [np.minimum((I * g.random(nnz) ** alpha).astype(np.int64), I - 1) for I in shape]
    vals = g.random(nnz) - 0.25
    U = [g.random((I, R)) - 0.5 for I in shape]
    for n in range(len(shape)):
        ref = orc.mttkrp_coo(tuple(subs), vals, shape, U, n)
        got = gpu_mttkrp_coo(K, dev, subs, vals, shape, U, n)
        assert relerr(got, ref) < TOL, (shape, n)


def test_sparse_edge_cases(K, dev, sparse_path):
    from sktensor import sptensor
    # one heavy row spanning many CTAs, the rest empty; plus a lone entry in the last row
    nnz = 300000
    g = np.random.default_rng(1)
    subs = [np.full(nnz, 7, dtype=np.int64), g.integers(0, 50, nnz), g.integers(0, 40, nnz)]
    subs[0][-1] = 99
    vals = g.random(nnz)
    U = [g.random((I, 16)) for I in (100, 50, 40)]
    for n in range(3):
        assert relerr(gpu_mttkrp_coo(K, dev, subs, vals, (100, 50, 40), U, n),
                      orc.mttkrp_coo(tuple(subs), vals, (100, 50, 40), U, n)) < TOL
    # empty tensor -> zeros of the right shape
    e = [np.zeros(0, dtype=np.int64)] * 3
    M = gpu_mttkrp_coo(K, dev, e, np.zeros(0), (4, 5, 6), U[:3], 0) if False else None
    S = sptensor((np.zeros(0, dtype=np.int64),) * 3, np.zeros(0), shape=(4, 5, 6))
    M = S.uttkrp([np.ones((4, 2)), np.ones((5, 2)), np.ones((6, 2))], 1)
    assert M.shape == (5, 2) and (M == 0).all()
    # out-of-range subscripts are rejected, not read
    with pytest.raises(ValueError):
        K.CooHandle([np.array([0, 9]), np.array([0, 1]), np.array([0, 1])], np.ones(2), (3, 3, 3))
    # determinism
    a = gpu_mttkrp_coo(K, dev, subs, vals, (100, 50, 40), U, 1)
    b = gpu_mttkrp_coo(K, dev, subs, vals, (100, 50, 40), U, 1)
    assert (a == b).all()


def test_sparse_method_contract(K, dev, golden):
    from sktensor import sptensor
    name = 'rand5'
    shape = golden.j['sp_cases'][name]['shape']
    subs = tuple(golden['mts_%s_sub%d' % (name, m)] for m in range(5))
    S = sptensor(subs, golden['mts_%s_vals' % name], shape)          # integer values (fixture)
    SU = S.uttkrp([np.zeros((s, 5)) for s in shape], mode=0)         # test_sptensor.py:141-149
    assert SU.shape == (25, 5) and SU.dtype == np.float64 and (SU == 0).all()
    U = golden.factors('mts_%s' % name, 5)
    for n in range(5):
        assert relerr(S.uttkrp(U, n), golden['mts_%s_M%d' % (name, n)]) < TOL
    assert abs(S.norm() - golden.j['mts_%s_norm' % name]) < 1e-12


def test_sparse_power_law_200k(K, dev, sparse_path):
    """BASELINE config 3 generator at 2e5 nnz: parity vs the oracle, all modes."""
    shape = (10000000, 1000000, 100000)
    nnz = 200000
    rng = np.random.default_rng(0)
    subs = [np.minimum((I * rng.random(nnz) ** 3.0).astype(np.int64), I - 1) for I in shape]
    vals = rng.random(nnz)
    U = [rng.random((I, 16)) for I in shape]
    coo = K.CooHandle(subs, vals, shape)
    Ud = [dev.to_device(u) for u in U]
    for n in range(3):
        got = dev.to_host(K.mttkrp_coo(coo, Ud, 16, n))
        ref = orc.mttkrp_coo(tuple(subs), vals, shape, U, n)
        assert relerr(got, ref) < TOL, n
        at = dev.to_host(K.mttkrp_coo(coo, Ud, 16, n, atomic=True))
        assert relerr(got, at) < TOL


# ------------------------------------------------------------------ dimension tree (dimtree.cu)
@pytest.mark.parametrize('shape,R', [
    ((9, 10, 11), 3), ((64, 48, 40), 32), ((33, 17, 20), 7), ((12, 10, 8, 14), 16), ((6, 5, 7, 4, 3), 4),
    ((300, 260, 36), 32), ((40, 24, 20, 36), 64),
])
def test_dimension_tree_matches_oracle(K, dev, shape, R, dense_path):
    """Every split point s: the partial contraction (dense kernel on the reshaped view; for a one-mode
    right group with L = T = 1 this is the direct-store path) followed by skt_kr_reduce reproduces
    dtensor.uttkrp (dtensor.py:163-167) for every mode of both groups."""
    rng = np.random.RandomState(hash((shape, R, 7)) % (2 ** 31))
    N = len(shape)
    X = rng.rand(*shape) - 0.4
    U = [rng.rand(d, R) - 0.5 for d in shape]
    Xd = dev.to_device(X)
    Ud = [dev.to_device(u) for u in U]
    ref = [orc.mttkrp_dense(X, U, n) for n in range(N)]
    for s in range(1, N):
        nA, nB = int(np.prod(shape[:s])), int(np.prod(shape[s:]))
        TA = K.mttkrp_dense(Xd, (nA,) + tuple(shape[s:]), [None] + Ud[s:], R, 0)
        TB = K.mttkrp_dense(Xd, tuple(shape[:s]) + (nB,), Ud[:s] + [None], R, s)
        for n in range(N):
            if n < s:
                got = TA if s == 1 else K.kr_reduce(TA, shape[:s], Ud[:s], R, n)
            else:
                got = TB if s == N - 1 else K.kr_reduce(TB, shape[s:], Ud[s:], R, n - s)
            assert relerr(dev.to_host(got), ref[n]) < TOL, (shape, R, s, n)


def test_dimension_tree_engine_equals_plain_sweeps():
    """cp_als with and without the dimension tree: same iteration count, fit within 1e-12."""
    import os
    from sktensor import cp_als, dtensor
    rng = np.random.RandomState(11)
    for shape, R in (((40, 36, 44), 8), ((20, 18, 16, 22), 6)):
        X = rng.rand(*shape)
        res = {}
        for flag in ('0', '1'):
            os.environ['SKT_NO_DIMTREE'] = flag
            np.random.seed(5)
            P, fit, itr, _ = cp_als(dtensor(X), R, init='random', max_iter=12, conv=0)
            res[flag] = (float(fit), itr, [np.asarray(u) for u in P.U])
        os.environ['SKT_NO_DIMTREE'] = '0'
        assert res['0'][1] == res['1'][1]
        assert abs(res['0'][0] - res['1'][0]) < 1e-12
        for a, b in zip(res['0'][2], res['1'][2]):
            assert relerr(a, b) < 1e-9


def test_empty_slabs_are_legal(K, dev):
    """A rank of a sharded run may own no rows (7 rows over 8 ranks; power-law cuts): every kernel the engine
    calls must accept zero-sized slabs and contribute exact zeros to the collectives."""
    R = 4
    X = dev.to_device(np.zeros((6, 5, 4, 0)))
    U = [dev.to_device(np.random.rand(6, R)), dev.to_device(np.random.rand(5, R)), dev.to_device(np.random.rand(4, R)),
         dev.to_device(np.zeros((0, R)))]
    for n in range(3):
        M = K.mttkrp_dense(X, (6, 5, 4, 0), U, R, n)
        assert M.shape[0] == (6, 5, 4)[n] and (dev.to_host(M) == 0).all()
    assert K.mttkrp_dense(X, (6, 5, 4, 0), U, R, 3).shape == (0, R)
    TA = K.mttkrp_dense(X, (30, 4, 0), [None, U[2], U[3]], R, 0)
    assert (dev.to_host(TA) == 0).all()
    TB = dev.to_device(np.zeros((4, 0, R)))
    assert (dev.to_host(K.kr_reduce(TB, (4, 0), [U[2], U[3]], R, 0)) == 0).all()
    assert K.kr_reduce(TB, (4, 0), [U[2], U[3]], R, 1).shape == (0, R)
    assert float(dev.to_host(K.sumsq(X))[0]) == 0.0
    assert (dev.to_host(K.gram(U[3])) == 0).all()
    P = dev.to_device(np.eye(R))
    for first in (True, False):
        Unew, colstat = K.solve_stats(dev.to_device(np.zeros((0, R))), P, first)
        c = dev.to_host(colstat)
        assert (c == 0).all() if first else (c < -1e300).all()
        lam, G, ip = K.normalise_gram(Unew, dev.to_device(np.ones(R)), first, Mip=dev.to_device(np.zeros((0, R))))
        assert (dev.to_host(G) == 0).all() and float(dev.to_host(ip)[0]) == 0.0
    coo = K.CooHandle([np.zeros(0, np.int64)] * 3, np.zeros(0), (5, 0, 3))
    Us = [dev.to_device(np.random.rand(5, R)), dev.to_device(np.zeros((0, R))), dev.to_device(np.random.rand(3, R))]
    assert (dev.to_host(K.mttkrp_coo(coo, Us, R, 0)) == 0).all()
    assert K.mttkrp_coo(coo, Us, R, 1).shape == (0, R)


@pytest.mark.parametrize('shape,nnz,R,prow', [
    ((100, 80, 60), 5000, 16, 8), ((1000, 50, 20), 30000, 3, 100), ((64, 64, 64), 70000, 32, 1),
    ((50, 40, 30, 20), 8000, 8, 7), ((20000, 300, 7), 60000, 64, 4096), ((300, 200), 4000, 4, 33),
    ((5, 200000, 3), 50000, 33, 50000),
])
def test_sparse_panelled_copy_vs_oracle(K, dev, shape, nnz, R, prow, sparse_path, monkeypatch):
    """Index-panel blocking (skt_coo_build sorts by (panel of the largest gathered mode, output subscript) and the
    kernel writes one slab of M per panel): forced on small tensors, every mode against the oracle; the atomic
    check kernel, <P,X> and the norm read the same panelled copy."""
    monkeypatch.setenv('SKT_COO_PANEL_ROWS', str(prow))
    monkeypatch.setenv('SKT_COO_PANEL_CAP', '0')
    g = np.random.default_rng(abs(hash((shape, nnz, R, prow))) % (2 ** 31))
    subs = [np.minimum((I * g.random(nnz) ** 2.0).astype(np.int64), I - 1) for I in shape]
    vals = g.random(nnz) - 0.25
    U = [g.random((I, R)) - 0.5 for I in shape]
    coo = K.CooHandle(subs, vals, shape)
    Ud = [dev.to_device(u) for u in U]
    for n in range(len(shape)):
        ref = orc.mttkrp_coo(tuple(subs), vals, shape, U, n)
        got = dev.to_host(K.mttkrp_coo(coo, Ud, R, n))
        assert relerr(got, ref) < TOL, (shape, n)
        assert relerr(dev.to_host(K.mttkrp_coo(coo, Ud, R, n, atomic=True)), ref) < TOL, (shape, n)
        again = dev.to_host(K.mttkrp_coo(coo, Ud, R, n))
        assert (got == again).all()
    lam = g.random(R)
    ip = float(dev.to_host(K.coo_innerprod(coo, Ud, dev.to_device(lam), R))[0])
    ref_ip = float(np.ravel(orc.kruskal_innerprod_coo(U, lam, tuple(subs), vals, shape))[0])
    assert abs(ip - ref_ip) < 1e-10 * max(abs(ref_ip), 1.0)
    assert abs(float(dev.to_host(K.coo_sumsq(coo))[0]) - float(np.sum(vals ** 2))) < 1e-10 * float(np.sum(vals ** 2))
    coo.close()

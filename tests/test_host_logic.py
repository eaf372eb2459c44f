"""Host-side behaviour of the drop-in package that needs no GPU: containers, index
bookkeeping, error vocabulary, and the reference's own CPU-runnable tests restated
(sktensor/tests/test_base.py, test_dtensor.py, test_sptensor.py, test_utils.py, test_pyutils.py)."""
import numpy as np
import pytest

import sktensor
from sktensor import dtensor, sptensor, ktensor, khatrirao, check_multiplication_dims
from sktensor.pyutils import from_to_without
from sktensor.sptensor import fromarray
from sktensor.utils import accum
from sktensor import _parallel


@pytest.fixture
def T():
    T = np.zeros((3, 4, 2))
    T[:, :, 0] = [[1, 4, 7, 10], [2, 5, 8, 11], [3, 6, 9, 12]]
    T[:, :, 1] = [[13, 16, 19, 22], [14, 17, 20, 23], [15, 18, 21, 24]]
    return T


@pytest.fixture
def rand5():
    # sktensor/tests/sptensor_rand_fixture.py:5-27
    shape = (25, 11, 18, 7, 2)
    np.random.seed(5)
    vals = np.random.randint(0, 100, 100)
    np.random.seed(5)
    subs = tuple(np.random.randint(0, shape[i], 100) for i in range(len(shape)))
    return subs, vals, shape


def test_public_names():
    for name in ('dtensor', 'sptensor', 'ktensor', 'unfolded_dtensor', 'unfolded_sptensor', 'cp_als',
                 'tucker_hooi', 'tucker_hosvd', 'khatrirao', 'nvecs', 'norm', 'ttm', 'ttv', 'unfold',
                 'transpose', 'accum', 'check_multiplication_dims', 'flipsign', 'innerprod'):
        assert hasattr(sktensor, name), name


def test_check_multiplication_dims():          # test_base.py:9-14
    assert ([1, 2] == check_multiplication_dims(0, 3, 2, without=True)).all()
    assert ([0, 2] == check_multiplication_dims(1, 3, 2, without=True)).all()
    assert ([0, 1] == check_multiplication_dims(2, 3, 2, without=True)).all()
    with pytest.raises(ValueError, match='Invalid dimensions'):
        check_multiplication_dims(5, 3, 2)
    with pytest.raises(ValueError, match='More multiplicants'):
        check_multiplication_dims([0], 3, 4, vidx=True)
    with pytest.raises(ValueError, match='Invalid number'):
        check_multiplication_dims([0], 3, 2, vidx=True)
    d, v = check_multiplication_dims([2, 0], 3, 2, vidx=True)
    assert list(d) == [0, 2] and list(v) == [1, 0]
    d, v = check_multiplication_dims([2, 0], 3, 3, vidx=True)
    assert list(d) == [0, 2] and list(v) == [0, 2]


def test_khatrirao():                          # test_base.py:17-42
    A = np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9]])
    B = np.array([[1, 4, 7], [2, 5, 8], [3, 6, 9]])
    C = np.array([[1, 8, 21], [2, 10, 24], [3, 12, 27], [4, 20, 42], [8, 25, 48], [12, 30, 54],
                  [7, 32, 63], [14, 40, 72], [21, 48, 81]])
    D = khatrirao((A, B))
    assert C.shape == D.shape and (C == D).all()
    rng = np.random.RandomState(0)
    Ms = tuple(rng.rand(n, 4) for n in (3, 5, 2))
    for rev in (False, True):
        want = np.zeros((30, 4))
        order = Ms[::-1] if rev else Ms
        for r in range(4):
            col = order[0][:, r]
            for M in order[1:]:
                col = np.kron(col, M[:, r])
            want[:, r] = col
        assert np.allclose(khatrirao(Ms, reverse=rev), want, rtol=1e-15)
    with pytest.raises(ValueError):
        khatrirao([A, B])


def test_dtensor_new():                        # test_dtensor.py:7-14
    A = np.random.randn(10, 23, 5)
    X = dtensor(A)
    assert A.ndim == X.ndim and A.shape == X.shape
    assert (A == X).all() and (X == A).all()
    assert np.shares_memory(A, X)              # zero-copy view (dtensor.py:48-50)


def test_dense_unfold_order(T):                # test_base.py:45-66, test_dtensor.py:17-38
    X = dtensor(T)
    I, J, K = T.shape
    X1, X2 = X[:, :, 0], X[:, :, 1]
    U = X.unfold(0)
    assert (3, 8) == U.shape
    for j in range(J):
        assert (U[:, j] == X1[:, j]).all() and (U[:, j + J] == X2[:, j]).all()
    U = X.unfold(1)
    assert (4, 6) == U.shape
    for i in range(I):
        assert (U[:, i] == X1[i, :]).all() and (U[:, i + I] == X2[i, :]).all()
    U = X.unfold(2)
    assert (2, 12) == U.shape
    for k in range(12):
        assert (U[:, k] == np.array([np.asarray(X1).flatten('F')[k], np.asarray(X2).flatten('F')[k]])).all()


def test_dtensor_fold_unfold():                # test_base.py:69-74
    X = dtensor(np.random.randn(10, 35, 3, 12))
    for i in range(4):
        assert (X == X.unfold(i).fold()).all()


def test_accum():                              # test_utils.py:5-20
    nvals, nsubs = accum((np.array([0, 1, 1, 2, 2, 2]), np.array([0, 1, 1, 1, 2, 2])),
                         np.array([1, 2, 3, 4, 5, 6]), with_subs=True)
    assert np.allclose(nvals, [1, 5, 4, 11])
    assert np.allclose(nsubs[0], [0, 1, 2, 2]) and np.allclose(nsubs[1], [0, 1, 1, 2])
    nvals, nsubs = accum((np.array([0, 0, 1]), np.array([0, 0, 1])), np.array([1, 2, 3]), with_subs=True)
    assert np.allclose(nvals, [3, 3]) and np.allclose(nsubs[0], [0, 1]) and np.allclose(nsubs[1], [0, 1])


def test_from_to_without():                    # test_pyutils.py:4-11
    frm, to, without = 2, 88, 47
    lst = list(range(frm, without)) + list(range(without + 1, to))
    assert lst == from_to_without(frm, to, without)
    assert lst[::-1] == from_to_without(frm, to, without, reverse=True)
    a, b = from_to_without(0, 5, 2, separate=True)
    assert a == [0, 1] and b == [3, 4]


def test_sptensor_init(rand5):                 # test_sptensor.py:22-54
    subs, vals, shape = rand5
    S = sptensor(subs, vals, shape)
    assert len(shape) == S.ndim and tuple(shape) == tuple(S.shape)
    S = sptensor(subs, vals)
    assert (np.array(subs).max(axis=1) + 1 == np.array(S.shape)).all()
    n = 20
    dsubs = tuple(np.arange(n) for _ in range(3))
    assert sptensor(dsubs, np.ones(n)).shape == (n, n, n)
    with pytest.raises(ValueError, match='tuple'):
        sptensor(np.random.randint(0, 10, 18).reshape(3, 3, 2), np.ones(10))
    with pytest.raises(ValueError, match='equal length'):
        sptensor(subs, np.ones(len(subs) + 1))
    with pytest.raises(ValueError, match='integers'):
        sptensor((np.array([0, 1], dtype=np.uint32), np.array([0, 1])), [1., 2.])
    # accumfun merges duplicates at construction (sptensor.py:80-84)
    S = sptensor((np.array([0, 0, 1]), np.array([2, 2, 0])), np.array([1., 2., 5.]), accumfun=np.sum)
    assert len(S.vals) == 2 and sorted(S.vals) == [3., 5.]


def test_sptensor_unfold_fold(rand5):          # test_sptensor.py:57-93
    subs, vals, shape = rand5
    Td = dtensor(np.zeros(shape, dtype=np.float32))
    Td[subs] = vals
    for i in range(len(shape)):
        rdims = [i]
        cdims = np.setdiff1d(range(len(shape)), rdims)[::-1]
        Md = Td.unfold(i)
        S = sptensor(subs, vals, shape, accumfun=lambda l: l[-1])
        for Ms in (S.unfold(rdims, cdims), S.unfold(rdims)):
            assert Md.shape == Ms.shape and np.allclose(Md, Ms.toarray())
        Ms = S.unfold(rdims, cdims, transp=True)
        assert Md.T.shape == Ms.shape and np.allclose(Md.T, Ms.toarray())
    S = sptensor(subs, vals, shape)
    for i in range(len(shape)):
        assert S.unfold([i]).fold() == S


def test_sptensor_ttm_needs_the_device(T):     # sparse TTM runs on the sparse MTTKRP kernel: no host route is left
    import torch
    S = sptensor(T.nonzero(), T.flatten(), T.shape)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            S.ttm(np.array([[1, 3, 5], [2, 4, 6]]), 0)
    with pytest.raises(ValueError):             # argument checks come first (dtensor.py:62 analogue)
        S.ttm(np.ones((2, 7)), 0)


def test_sptensor_getitem_sub_toarray():       # test_sptensor.py:152-183
    subs = (np.array([0, 1, 0, 5, 7, 8]), np.array([2, 0, 4, 5, 3, 9]), np.array([0, 1, 2, 2, 1, 0]))
    S = sptensor(subs, np.array([1, 2, 3, 4, 5, 6]), shape=[10, 10, 3])
    for idx, want in (((1, 1, 1), 0), ((0, 2, 0), 1), ((1, 0, 1), 2), ((0, 4, 2), 3), ((5, 5, 2), 4),
                      ((7, 3, 1), 5), ((8, 9, 0), 6)):
        assert want == S[idx]
    S = sptensor((np.array([0, 1, 0]), np.array([2, 0, 2]), np.array([0, 1, 2])), np.array([1, 2, 3]), shape=[3, 3, 3])
    D = np.arange(27).reshape(3, 3, 3)
    R = S - D
    for i in range(3):
        for j in range(3):
            for k in range(3):
                assert S[i, j, k] - D[i, j, k] == R[i, j, k]
    A = np.random.RandomState(1).rand(4, 3, 5) * (np.random.RandomState(2).rand(4, 3, 5) > 0.5)
    assert (fromarray(A).toarray() == A).all()
    c = fromarray(A).concatenate((fromarray(A), fromarray(A)), axis=1)
    assert c.shape == (4, 6, 5) and np.allclose(c.toarray(), np.concatenate((A, A), axis=1))
    assert fromarray(A).transpose((2, 0, 1)).shape == (5, 4, 3)


def test_ktensor_container():                  # ktensor.py:59-68, test_ktensor.py:5-14 (fixed for py3)
    U = [np.random.randn(s, 5) for s in (5, 27, 3, 13)]
    K = ktensor(U)
    assert K.shape == (5, 27, 3, 13) and K.ndim == 4 and K.rank == 5 and (K.lmbda == 1).all()
    assert K.U is U                                           # aliased, not copied
    v = K.tovec()
    assert sum(s * 5 for s in K.shape) == len(v.v)
    assert K == v.toktensor()
    with pytest.raises(ValueError, match='Dimension mismatch'):
        ktensor([np.ones((3, 2)), np.ones((4, 3))])


def test_cp_als_argument_errors_before_any_device_work():
    X = dtensor(np.zeros((3, 4, 2)))
    with pytest.raises(ValueError, match='Unknown keywords'):
        sktensor.cp_als(X, 2, maxiter=3)
    with pytest.raises(ValueError, match='Unknown keywords'):
        sktensor.tucker_hooi(X, [2, 2, 2], max_iter=3)


def test_no_cpu_fallback():
    """Without a CUDA device the path methods must raise, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    X = dtensor(np.random.rand(3, 4, 2))
    U = [np.random.rand(s, 2) for s in X.shape]
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        X.uttkrp(U, 0)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        fromarray(np.asarray(X)).uttkrp(U, 1)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        np.random.seed(0)
        sktensor.cp_als(X, 2, init='random')
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        X.norm()


def test_product_never_imports_the_oracle():
    import os
    from conftest import PKG, ROOT
    for base, _, files in os.walk(PKG):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(base, f)).read()
                assert 'oracle' not in src.lower().replace('tests/fake_backend', ''), os.path.join(base, f)


def test_partition_helpers():
    assert list(_parallel.split_rows(10, 3)) == [0, 4, 7, 10]
    assert list(_parallel.split_rows(2, 4)) == [0, 1, 2, 2, 2]
    rng = np.random.default_rng(0)
    idx = np.minimum((1000 * rng.random(50000) ** 3).astype(np.int64), 999)
    for world in (2, 4, 8):
        offs = _parallel.split_nnz(idx, 1000, world)
        assert offs[0] == 0 and offs[-1] == 1000 and (np.diff(offs) >= 0).all()
        counts = [np.count_nonzero((idx >= offs[r]) & (idx < offs[r + 1])) for r in range(world)]
        assert sum(counts) == len(idx)
        heaviest_row = np.bincount(idx).max()
        assert max(counts) <= len(idx) / world + heaviest_row
    assert _parallel.shard_mode((4, 9, 9, 2)) == 1


def test_tree_split():
    from sktensor._parallel import tree_split
    assert tree_split((512, 512, 512)) == 2          # tie -> last-mode MTTKRP is the second pass
    assert tree_split((160, 160, 160, 160)) == 2
    assert tree_split((10, 11)) == 1
    assert tree_split((1000, 10, 10)) == 1
    assert tree_split((10, 10, 1000)) == 2
    assert tree_split((4, 5, 6, 7, 8)) in (2, 3)


def test_shard_mode_prefers_middle_modes_for_dense_ties():
    from sktensor._parallel import shard_mode
    assert shard_mode((160, 160, 160, 160)) == 0
    assert shard_mode((160, 160, 160, 160), dense=True) == 1
    assert shard_mode((512, 512, 512), dense=True) == 1
    assert shard_mode((1024, 512, 512), dense=True) == 0
    assert shard_mode((10, 11, 8), dense=True) == 1
    assert shard_mode((10, 8, 11), dense=True) == 2
    assert shard_mode((7, 7), dense=True) == 0


def test_local_shard_on_a_single_rank_is_the_whole_tensor():
    """A LocalShard handed to a single-process engine must not slice the given factors (regression: bench cfg5 at N=1)."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from fake_backend import NumpyBackend
    from sktensor import _parallel, dtensor
    rng = np.random.RandomState(0)
    X = rng.rand(6, 7, 5)
    U = [None, rng.rand(7, 3), rng.rand(5, 3)]
    shard = _parallel.LocalShard(dtensor(X), 1, X.shape, np.array([0, 7]))
    eng = _parallel.CpAlsEngine(shard, 3, list(U), backend=NumpyBackend())
    assert eng.s == -1 and tuple(eng.Ud[1].shape) == (7, 3)
    eng.sweep(first_iter=True)
    ref = _parallel.CpAlsEngine(dtensor(X), 3, list(U), backend=NumpyBackend())
    ref.sweep(first_iter=True)
    assert float(eng.fitout[0]) == float(ref.fitout[0])

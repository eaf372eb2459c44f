"""Pin the CPU oracle (oracle/sktensor_oracle.py) against vectors produced by the
unmodified reference (oracle/make_golden.py) and against the reference's own
known-answer tests.  CPU only."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, relerr
from oracle import sktensor_oracle as orc


# ---- the reference's own known-answer tests, restated against the oracle ------------
def test_khatrirao_known_answer():
    # sktensor/tests/test_base.py:17-42
    A = np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9]])
    B = np.array([[1, 4, 7], [2, 5, 8], [3, 6, 9]])
    C = np.array([[1, 8, 21], [2, 10, 24], [3, 12, 27], [4, 20, 42], [8, 25, 48],
                  [12, 30, 54], [7, 32, 63], [14, 40, 72], [21, 48, 81]])
    D = orc.khatrirao((A, B))
    assert D.shape == C.shape and (D == C).all()
    with pytest.raises(ValueError):
        orc.khatrirao([A, B])
    with pytest.raises(ValueError):
        orc.khatrirao((A, B[:, :2]))


def test_unfold_column_order(golden):
    # sktensor/tests/test_base.py:45-66 (Kolda-Bader order)
    T = golden['kat1_T']
    I, J, K = T.shape
    X1, X2 = T[:, :, 0], T[:, :, 1]
    U = orc.unfold(T, 0)
    assert U.shape == (3, 8)
    for j in range(J):
        assert (U[:, j] == X1[:, j]).all() and (U[:, j + J] == X2[:, j]).all()
    U = orc.unfold(T, 1)
    assert U.shape == (4, 6)
    for i in range(I):
        assert (U[:, i] == X1[i, :]).all() and (U[:, i + I] == X2[i, :]).all()
    U = orc.unfold(T, 2)
    assert U.shape == (2, 12)
    for k in range(12):
        assert (U[:, k] == np.array([X1.flatten('F')[k], X2.flatten('F')[k]])).all()


def test_fold_unfold_identity():
    # sktensor/tests/test_base.py:69-74
    X = np.random.RandomState(0).randn(10, 35, 3, 12)
    for m in range(4):
        assert (orc.fold(orc.unfold(X, m), m, X.shape) == X).all()


def test_ttm_known_answer(golden):
    # sktensor/tests/test_base.py:77-81 with ttm_fixture.py:14-23
    Y = np.zeros((2, 4, 2))
    Y[:, :, 0] = [[22, 49, 76, 103], [28, 64, 100, 136]]
    Y[:, :, 1] = [[130, 157, 184, 211], [172, 208, 244, 280]]
    Y2 = orc.ttm_dense(golden['kat1_T'], golden['ttm_V'], 0)
    assert Y2.shape == (2, 4, 2) and (Y2 == Y).all()
    assert (golden['ttm_Y0'] == Y).all()


def test_sparse_ttv_known_answer():
    # sktensor/tests/test_sptensor.py:121-133
    subs = (np.array([0, 1, 0, 5, 7, 8]), np.array([2, 0, 4, 5, 3, 9]), np.array([0, 1, 2, 2, 1, 0]))
    vals = np.array([1, 2, 3, 4, 5, 6.1])
    # contract two of three modes -> vector, compare with a dense einsum of the same tensor
    w0, w1 = np.arange(1., 11.), np.arange(1., 13.)
    rows, sums = orc.ttv_coo(subs, vals, (10, 12, 3), (w0, w1), [0, 1])
    dense = np.einsum('ijk,i,j->k', orc.coo_toarray(subs, vals, (10, 12, 3)), w0, w1)
    out = np.zeros(3)
    out[rows] = sums
    assert np.allclose(out, dense, rtol=1e-14)


def test_sparse_mttkrp_zero_factor_shape(golden):
    # sktensor/tests/test_sptensor.py:141-149 -> (25, 5)
    name = 'rand5'
    shape = golden.j['sp_cases'][name]['shape']
    subs = tuple(golden['mts_%s_sub%d' % (name, m)] for m in range(len(shape)))
    Uz = [np.zeros((s, 5)) for s in shape]
    M = orc.mttkrp_coo(subs, golden['mts_%s_vals' % name], shape, Uz, 0)
    assert M.shape == (25, 5) == tuple(golden.j['rand5_zero_shape'])


# ---- golden vectors from the unmodified reference ------------------------------------
def test_kat1_exact(golden):
    U = golden.factors('kat1', 3)
    T = golden['kat1_T']
    exp = {0: [[1048, 1800], [1112, 1920], [1176, 2040]],
           1: [[412, 744], [520, 960], [628, 1176], [736, 1392]],
           2: [[1270, 2000], [2998, 4880]]}
    for n in range(3):
        assert (golden['kat1_M%d' % n] == np.array(exp[n])).all()
        assert (orc.mttkrp_dense(T, U, n) == np.array(exp[n])).all()
        s, v, shp = orc.coo_fromarray(T)
        assert (orc.mttkrp_coo(s, v, shp, U, n) == np.array(exp[n])).all()


def test_mttkrp_dense_golden(golden):
    for name, c in golden.j['dense_cases'].items():
        N = len(c['shape'])
        X = golden['mtd_%s_X' % name]
        U = golden.factors('mtd_%s' % name, N)
        for n in range(N):
            Uin = list(U)
            Uin[n] = None
            M = orc.mttkrp_dense(X, Uin, n)
            assert relerr(M, golden['mtd_%s_M%d' % (name, n)]) < 1e-14, (name, n)
            # independent einsum statement of the MTTKRP contract (core.py:145-183)
            letters = 'abcdefg'[:N]
            expr = letters + ',' + ','.join(letters[m] + 'r' for m in range(N) if m != n) + '->' + letters[n] + 'r'
            E = np.einsum(expr, X, *[U[m] for m in range(N) if m != n])
            assert relerr(E, M) < 1e-12


def test_mttkrp_coo_golden(golden):
    for name, c in golden.j['sp_cases'].items():
        N = len(c['shape'])
        subs = tuple(golden['mts_%s_sub%d' % (name, m)] for m in range(N))
        vals = golden['mts_%s_vals' % name]
        U = golden.factors('mts_%s' % name, N)
        assert abs(orc.norm_coo(vals) - golden.j['mts_%s_norm' % name]) < 1e-12
        for n in range(N):
            Uin = list(U)
            Uin[n] = None
            M = orc.mttkrp_coo(subs, vals, c['shape'], Uin, n)
            G = golden['mts_%s_M%d' % (name, n)]
            assert M.dtype == np.float64 and M.shape == G.shape
            assert relerr(M, G) < 1e-14, (name, n)


def test_solve_normalise_golden(golden):
    for name in golden.j['solve_cases']:
        M, Y = golden['slv_%s_M' % name], golden['slv_%s_Y' % name]
        F = [None, golden['slv_%s_F0' % name], golden['slv_%s_F1' % name]]
        assert relerr(orc.gram_hadamard(F, 0, M.shape[1]), Y) < 1e-15
        for itr in (0, 1):
            U, lam = orc.solve_normalise(M, Y, itr)
            assert relerr(U, golden['slv_%s_U%d' % (name, itr)]) < 1e-12, name
            assert relerr(lam, golden['slv_%s_lam%d' % (name, itr)]) < 1e-12, name


def test_kruskal_golden(golden):
    U = golden.factors('kt', 3)
    lam = golden['kt_lmbda']
    assert abs(orc.kruskal_norm(U, lam) - golden.j['kt_norm']) < 1e-12
    assert abs(orc.kruskal_innerprod_dense(U, lam, golden['kt_X']) - golden.j['kt_innerprod_dense']) < 1e-11
    s, v, shp = orc.coo_fromarray(golden['kt_Xs'])
    assert abs(orc.kruskal_innerprod_coo(U, lam, s, v, shp) - golden.j['kt_innerprod_sparse']) < 1e-11
    assert relerr(orc.kruskal_toarray(U, lam), golden['kt_toarray']) < 1e-14


def _check_cp(golden, key, res, N, tol=1e-9):
    U, lam, fit, itr, _ = res
    assert itr == golden.j[key]['itr'], key
    assert abs(float(np.ravel(fit)[0]) - golden.j[key]['fit']) < 1e-10, key
    assert relerr(lam, golden['%s_lmbda' % key]) < tol
    for m in range(N):
        assert relerr(U[m], golden['%s_U%d' % (key, m)]) < tol, (key, m)


def test_cp_als_cfg1_golden(golden):
    X = golden['cfg1_X']
    assert abs(orc.norm_dense(X) - golden.j['cfg1_normX']) < 1e-13
    for seed in (42, 7, 0):
        np.random.seed(seed)
        _check_cp(golden, 'cp_cfg1_s%d' % seed, orc.cp_als(X, 3, init='random'), 3)
    _check_cp(golden, 'cp_cfg1_nvecs', orc.cp_als(X, 3), 3)
    np.random.seed(42)
    _check_cp(golden, 'cp_cfg1_s42_it5', orc.cp_als(X, 3, init='random', max_iter=5, conv=0), 3)
    np.random.seed(42)
    res = orc.cp_als(X, 3, init='random', fit_method=None, max_iter=10)
    assert res[2] == 9 and res[3] == 9
    np.random.seed(42)
    trace = []
    orc.cp_als(X, 3, init='random', trace=trace)
    ref = golden.j['cp_cfg1_s42_trace']
    assert np.allclose(trace, ref[:len(trace)], rtol=0, atol=1e-11)
    np.random.seed(42)
    _check_cp(golden, 'cp_cfg1_s42_F', orc.cp_als(np.asfortranarray(X), 3, init='random'), 3)


def test_cp_als_sparse_and_singular_golden(golden):
    X = golden['cfg1_X']
    S = orc.coo_fromarray(X * (X > 0.7))
    assert len(S[1]) == golden.j['kat3_nnz']
    assert abs(orc.norm_coo(S[1]) - golden.j['kat3_norm']) < 1e-13
    np.random.seed(42)
    _check_cp(golden, 'cp_kat3_s42', orc.cp_als(S, 3, init='random'), 3)
    np.random.seed(42)
    _check_cp(golden, 'cp_kat3_dense_s42', orc.cp_als(orc.coo_toarray(*S), 3, init='random'), 3)
    init = [None, golden['kat5_init1'].copy(), golden['kat5_init2'].copy()]
    res = orc.cp_als(X, 3, init=init, max_iter=50)
    _check_cp(golden, 'cp_kat5', res, 3, tol=1e-6)
    assert init[0] is not None and init[1] is res[0][1]      # caller's list mutated (cp.py:195-196,154)
    np.random.seed(3)
    _check_cp(golden, 'cp_d4_s3', orc.cp_als(golden['cp4_X'], 4, init='random', max_iter=30), 4)
    np.random.seed(1)
    _check_cp(golden, 'cp_planted_s1', orc.cp_als(golden['cppl_X'], 3, init='random', max_iter=60), 3, tol=1e-7)
    subs = tuple(golden['cpsp_sub%d' % m] for m in range(3))
    np.random.seed(1)
    _check_cp(golden, 'cp_sp_s1', orc.cp_als((subs, golden['cpsp_vals'], (60, 50, 40)), 5,
                                             init='random', max_iter=25), 3)


def test_ttm_nvecs_hooi_golden(golden):
    X = golden['ttm_X']
    V = [golden['ttm_V%d' % m] for m in range(3)]
    for m in range(3):
        assert relerr(orc.ttm_dense(X, V[m], m, transp=True), golden['ttm_single%d' % m]) < 1e-14
        assert relerr(orc.ttm_chain(X, V, m, transp=True), golden['ttm_without%d' % m]) < 1e-14
        assert relerr(orc.nvecs_dense(X, m, 3), golden['nvecs_%d' % m]) < 1e-10
    assert relerr(orc.ttm_chain(X, V, None, transp=True), golden['ttm_all']) < 1e-14
    core, U, fit, _ = orc.hooi(golden['cfg1_X'], [3, 4, 2], init='nvecs')
    assert abs(np.linalg.norm(core) - golden.j['hooi_cfg1']['core_norm']) < 1e-10
    assert abs(fit - golden.j['hooi_cfg1']['fit']) < 1e-10
    assert relerr(core, golden['hooi_cfg1_core']) < 1e-8
    for m in range(3):
        assert relerr(U[m], golden['hooi_cfg1_U%d' % m]) < 1e-8
    np.random.seed(3)
    core, U, fit, _ = orc.hooi(golden['cfg1_X'], [3, 4, 2], init='random')
    assert abs(fit - golden.j['hooi_cfg1_rand3']['fit']) < 1e-10
    core, U, fit, itr = orc.hooi(golden['hooi_b_X'], [4, 5, 3], init='nvecs', maxIter=6, conv=0)
    assert itr == 5 and relerr(core, golden['hooi_b_core']) < 1e-8


# ---- the committed fixtures are reproducible from the reference (this container only) --
@pytest.mark.skipif(not os.path.isdir('/root/reference/sktensor'), reason='reference tree not present')
def test_golden_reproducible_from_reference(tmp_path, golden):
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE='1', PYTHONWARNINGS='ignore')
    subprocess.run([sys.executable, os.path.join(ROOT, 'oracle', 'make_golden.py'), '--out', str(tmp_path)],
                   check=True, env=env, capture_output=True, timeout=600)
    new = np.load(os.path.join(str(tmp_path), 'golden.npz'))
    assert sorted(new.files) == sorted(golden.a.files)
    for k in new.files:
        assert new[k].shape == golden.a[k].shape, k
        assert np.allclose(new[k], golden.a[k], rtol=1e-12, atol=1e-13, equal_nan=True), k


def test_rescal_oracle_matches_reference(golden):
    """RESCAL-ALS restatement (rescal.py:166-271) against the unmodified reference's factors and iteration counts."""
    for name, c in golden.j['rescal_cases'].items():
        X = [golden['rescal_%s_X%d' % (name, i)] for i in range(c['k'])]
        np.random.seed(c['seed'])
        A0 = np.random.rand(c['n'], c['rank'])
        A, R, fval, itr = orc.rescal_als(X, c['rank'], A0, lambda_A=c['lambda_A'], lambda_R=c['lambda_R'])
        assert itr == c['itr'], name
        assert np.max(np.abs(A - golden['rescal_%s_A' % name])) < 1e-9
        assert np.max(np.abs(np.array(R) - golden['rescal_%s_R' % name])) < 1e-8
        assert c['f'] == 0                       # the reference returns its never-assigned f (rescal.py:170,208)

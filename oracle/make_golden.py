"""Generate ``tests/golden/golden.npz`` + ``golden.json`` from the UNMODIFIED reference.

Run in this container only (``/root/reference`` is not on the GPU box):

    python oracle/make_golden.py            # writes tests/golden/
    python oracle/make_golden.py --out DIR  # elsewhere (used by the reproducibility test)

The reference is imported through ``oracle/ref_shim.py`` (aliases only, no
arithmetic touched).  Every vector stores its inputs as well as the reference's
outputs so the GPU-box tests need neither the reference nor a matching RNG.
Test infrastructure; not imported by the product.
"""
import argparse
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref = ref_shim.import_reference()
from sktensor import dtensor, sptensor, ktensor, cp_als, tucker_hooi  # noqa: E402
from sktensor.core import khatrirao, nvecs, ttm  # noqa: E402
from sktensor.sptensor import fromarray  # noqa: E402

A = {}   # arrays -> npz
J = {}   # scalars / small lists -> json


def put(name, arr):
    A[name] = np.asarray(arr)


def powerlaw_coo(rng, shape, nnz, alpha=3.0):
    """SURVEY 8(d) cfg3 generator: bounded power law, duplicates kept."""
    subs = tuple(np.minimum((I * rng.random(nnz) ** alpha).astype(np.int64), I - 1) for I in shape)
    return subs, rng.random(nnz)


# ---------------------------------------------------------------- KAT-1 (exact integers)
T = np.zeros((3, 4, 2))
T[:, :, 0] = [[1, 4, 7, 10], [2, 5, 8, 11], [3, 6, 9, 12]]
T[:, :, 1] = [[13, 16, 19, 22], [14, 17, 20, 23], [15, 18, 21, 24]]
Ua = np.arange(1, 7, dtype=float).reshape(3, 2)
Ub = np.arange(1, 9, dtype=float).reshape(4, 2)
Uc = np.array([[1., 2.], [3., 4.]])
put('kat1_T', T)
for m, Um in enumerate((Ua, Ub, Uc)):
    put('kat1_U%d' % m, Um)
for n in range(3):
    Md = np.asarray(dtensor(T).uttkrp([Ua, Ub, Uc], n))
    Ms = np.asarray(fromarray(T).uttkrp([Ua, Ub, Uc], n))
    assert (Md == Ms).all()
    put('kat1_M%d' % n, Md)

# ---------------------------------------------------------------- dense MTTKRP, assorted shapes
rng = np.random.RandomState(1234)
dense_cases = {
    'd3a': ((5, 6, 7), 4), 'd4a': ((4, 3, 5, 6), 3), 'd5a': ((3, 4, 2, 5, 3), 2),
    'cfg1': ((10, 11, 8), 3), 'd3b': ((33, 17, 66), 32), 'd4b': ((9, 16, 6, 10), 64),
    'd2a': ((7, 130), 5), 'd3odd': ((13, 7, 5), 6), 'd3c': ((130, 3, 40), 16),
}
J['dense_cases'] = {k: {'shape': list(v[0]), 'rank': v[1]} for k, v in dense_cases.items()}
for name, (shape, R) in dense_cases.items():
    X = rng.rand(*shape)
    U = [rng.rand(s, R) for s in shape]
    put('mtd_%s_X' % name, X)
    for m, Um in enumerate(U):
        put('mtd_%s_U%d' % (name, m), Um)
    for n in range(len(shape)):
        Uin = list(U)
        Uin[n] = None                      # U[mode] may be None (cp.py:194,143)
        put('mtd_%s_M%d' % (name, n), np.asarray(dtensor(X).uttkrp(Uin, n)))
    # layout independence (F-order input, Appendix A.15)
    Mf = np.asarray(dtensor(np.asfortranarray(X)).uttkrp(U, 0))
    assert np.allclose(Mf, A['mtd_%s_M0' % name], rtol=1e-13, atol=0)

# ---------------------------------------------------------------- sparse MTTKRP
sp_cases = {}
# the reference's own random fixture (tests/sptensor_rand_fixture.py:5-27): ints, duplicates possible
np.random.seed(5)
shape = (25, 11, 18, 7, 2)
vals = np.random.randint(0, 100, 100)
np.random.seed(5)
subs = tuple(np.random.randint(0, shape[i], 100) for i in range(len(shape)))
sp_cases['rand5'] = (subs, vals, shape, 5)
# the reference's 6-nnz fixture (tests/sptensor_fixture.py:6-21)
sp_cases['six'] = ((np.array([0, 1, 0, 5, 7, 8]), np.array([2, 0, 4, 5, 3, 9]), np.array([0, 1, 2, 2, 1, 0])),
                   np.array([1, 2, 3, 4, 5, 6.1]), (10, 12, 3), 4)
# KAT-6: duplicates with integer values
sp_cases['dup'] = ((np.array([0, 0, 1, 2, 2]), np.array([1, 1, 0, 2, 2]), np.array([0, 0, 1, 1, 1])),
                   np.array([1, 2, 3, 4, 5]), (3, 3, 2), 2)
g = np.random.default_rng(0)
s, v = powerlaw_coo(g, (2000, 300, 50), 20000)
sp_cases['pl3'] = (s, v, (2000, 300, 50), 16)
s, v = powerlaw_coo(g, (40, 30, 20, 10), 3000)
sp_cases['pl4'] = (s, v, (40, 30, 20, 10), 7)
s, v = powerlaw_coo(g, (5000, 64, 9), 4000, alpha=1.0)   # many empty rows
sp_cases['sparse_rows'] = (s, v, (5000, 64, 9), 8)
J['sp_cases'] = {k: {'shape': list(c[2]), 'rank': c[3], 'nnz': int(len(c[1]))} for k, c in sp_cases.items()}
for name, (subs, vals, shape, R) in sp_cases.items():
    if name == 'dup':
        U = [np.arange(1, 1 + 2 * I, dtype=float).reshape(I, 2) for I in shape]
    else:
        U = [g.random((I, R)) for I in shape]
    S = sptensor(tuple(subs), vals, shape=shape)
    for m in range(len(shape)):
        put('mts_%s_sub%d' % (name, m), np.asarray(subs[m], dtype=np.int64))
        put('mts_%s_U%d' % (name, m), U[m])
    put('mts_%s_vals' % name, vals)
    J['mts_%s_norm' % name] = float(S.norm())
    for n in range(len(shape)):
        Uin = list(U)
        Uin[n] = None
        put('mts_%s_M%d' % (name, n), S.uttkrp(Uin, n))
assert (A['mts_dup_M0'] == np.array([[9, 24], [9, 24], [135, 216]])).all()

# the reference's shape-only MTTKRP test (test_sptensor.py:141-149): all-zero factors -> (25, 5)
subs, vals, shape, _ = sp_cases['rand5']
S = sptensor(subs, vals, shape=shape)
Uz = [np.zeros((s_, 5)) for s_ in shape]
J['rand5_zero_shape'] = list(S.uttkrp(Uz, 0).shape)

# ---------------------------------------------------------------- sparse ttv known answer is in tests (test_sptensor.py:121-133)

# ---------------------------------------------------------------- normal-equation solve + normalise (cp.py:144-154)
from scipy.linalg import pinv  # noqa: E402
solve_cases = {}
for name, (I, R) in {'s3': (10, 3), 's16': (50, 16), 's32': (64, 32), 's64': (80, 64)}.items():
    Us = [rng.rand(40, R) for _ in range(2)]
    solve_cases[name] = (rng.rand(I, R) * 10 - 2, Us)
Us = [rng.rand(30, 4) for _ in range(2)]
for u in Us:
    u[:, 3] = u[:, 2]                       # collinear columns -> singular Y (Appendix A.17)
solve_cases['sing'] = (rng.rand(12, 4), Us)
Us = [rng.rand(30, 5) for _ in range(2)]
Us[0][:, 1] = 0                             # zero column -> rank-deficient Y
solve_cases['zerocol'] = (rng.rand(9, 5), Us)
J['solve_cases'] = list(solve_cases)
for name, (M, Us) in solve_cases.items():
    R = M.shape[1]
    Y = np.ones((R, R))
    for u in Us:
        Y = Y * np.dot(u.T, u)
    Unew = M.dot(pinv(Y))
    put('slv_%s_M' % name, M)
    put('slv_%s_Y' % name, Y)
    for k, u in enumerate(Us):
        put('slv_%s_F%d' % (name, k), u)
    put('slv_%s_Unew' % name, Unew)
    l0 = np.sqrt((Unew ** 2).sum(axis=0))
    put('slv_%s_lam0' % name, l0)
    put('slv_%s_U0' % name, Unew / l0)
    l1 = Unew.max(axis=0)
    l1[l1 < 1] = 1
    put('slv_%s_lam1' % name, l1)
    put('slv_%s_U1' % name, Unew / l1)

# ---------------------------------------------------------------- Kruskal norm / innerprod / toarray
Uk = [rng.rand(s_, 4) for s_ in (6, 5, 7)]
lk = rng.rand(4) + 0.5
Xk = rng.rand(6, 5, 7)
K = ktensor(Uk, lk)
for m in range(3):
    put('kt_U%d' % m, Uk[m])
put('kt_lmbda', lk)
put('kt_X', Xk)
put('kt_toarray', K.toarray())
J['kt_norm'] = float(K.norm())
J['kt_innerprod_dense'] = float(K.innerprod(dtensor(Xk)))
Sk = fromarray(Xk * (Xk > 0.6))
J['kt_innerprod_sparse'] = float(K.innerprod(Sk))
put('kt_Xs', Sk.toarray())

# ---------------------------------------------------------------- CP-ALS (KAT-2/3/5; cp.py:46-171)
np.random.seed(0)
X1 = np.random.rand(10, 11, 8)
put('cfg1_X', X1)
J['cfg1_normX'] = float(dtensor(X1).norm())


def run_cp(key, Xt, rank, seed=None, **kw):
    if seed is not None:
        np.random.seed(seed)
    P, fit, itr, _ = cp_als(Xt, rank, **kw)
    for m, Um in enumerate(P.U):
        put('%s_U%d' % (key, m), Um)
    put('%s_lmbda' % key, P.lmbda)
    J[key] = {'fit': float(fit), 'itr': int(itr)}
    return P, float(fit), int(itr)


for seed in (42, 7, 0):
    run_cp('cp_cfg1_s%d' % seed, dtensor(X1), 3, seed, init='random')
run_cp('cp_cfg1_nvecs', dtensor(X1), 3)
run_cp('cp_cfg1_s42_it5', dtensor(X1), 3, 42, init='random', max_iter=5, conv=0)
run_cp('cp_cfg1_nofit', dtensor(X1), 3, 42, init='random', fit_method=None, max_iter=10)
# per-iteration fit trace for seed 42 (max_iter=k, conv=0 replays the first k iterations)
trace = []
for k in range(1, 23):
    np.random.seed(42)
    _, fit, _, _ = cp_als(dtensor(X1), 3, init='random', max_iter=k, conv=0)
    trace.append(float(fit))
J['cp_cfg1_s42_trace'] = trace
# F-order input reproduces the same run (Appendix A.15)
_, fitF, itrF = run_cp('cp_cfg1_s42_F', dtensor(np.asfortranarray(X1)), 3, 42, init='random')
assert itrF == J['cp_cfg1_s42']['itr']

# KAT-3 sparse
S3 = fromarray(X1 * (X1 > 0.7))
J['kat3_nnz'] = int(len(S3.vals))
J['kat3_norm'] = float(S3.norm())
run_cp('cp_kat3_s42', S3, 3, 42, init='random')
run_cp('cp_kat3_dense_s42', dtensor(S3.toarray()), 3, 42, init='random')
run_cp('cp_kat3_nvecs', S3, 3)

# KAT-5 singular normal equations
np.random.seed(5)
init = [None] + [np.random.rand(s_, 3) for s_ in (11, 8)]
for u in init[1:]:
    u[:, 2] = u[:, 1]
put('kat5_init1', init[1].copy())
put('kat5_init2', init[2].copy())
P5, _, _ = run_cp('cp_kat5', dtensor(X1), 3, None, init=init, max_iter=50)
J['kat5_init_mutated'] = bool(init[0] is not None and init[1] is P5.U[1])

# 4-way dense + planted low-rank + sparse 4-way
X4 = rng.rand(6, 5, 4, 7)
put('cp4_X', X4)
run_cp('cp_d4_s3', dtensor(X4), 4, 3, init='random', max_iter=30)
Upl = [rng.rand(s_, 3) for s_ in (12, 9, 10)]
Xpl = ktensor(Upl).toarray() + 0.01 * rng.randn(12, 9, 10)
put('cppl_X', Xpl)
run_cp('cp_planted_s1', dtensor(Xpl), 3, 1, init='random', max_iter=60)
s, v = powerlaw_coo(np.random.default_rng(7), (60, 50, 40), 5000)
for m in range(3):
    put('cpsp_sub%d' % m, s[m])
put('cpsp_vals', v)
run_cp('cp_sp_s1', sptensor(s, v, shape=(60, 50, 40)), 5, 1, init='random', max_iter=25)

# ---------------------------------------------------------------- TTM / nvecs / HOOI (KAT-4)
Vt = np.array([[1., 3., 5.], [2., 4., 6.]])
put('ttm_V', Vt)
put('ttm_Y0', np.asarray(dtensor(T).ttm(Vt, 0)))
Xt = rng.rand(7, 6, 5)
Vs = [rng.rand(7, 3), rng.rand(6, 4), rng.rand(5, 2)]
put('ttm_X', Xt)
for m in range(3):
    put('ttm_V%d' % m, Vs[m])
    put('ttm_single%d' % m, np.asarray(dtensor(Xt).ttm(Vs[m], m, transp=True)))
    put('ttm_without%d' % m, np.asarray(ttm(dtensor(Xt), Vs, m, transp=True, without=True)))
    put('nvecs_%d' % m, nvecs(dtensor(Xt), m, 3))
put('ttm_all', np.asarray(dtensor(Xt).ttm(Vs, transp=True)))

core, Uh = tucker_hooi(dtensor(X1), [3, 4, 2], init='nvecs')
put('hooi_cfg1_core', np.asarray(core))
for m in range(3):
    put('hooi_cfg1_U%d' % m, Uh[m])
normres = np.sqrt(np.linalg.norm(X1) ** 2 - np.linalg.norm(core) ** 2)
J['hooi_cfg1'] = {'core_norm': float(np.linalg.norm(core)), 'fit': float(1 - normres / np.linalg.norm(X1))}
np.random.seed(3)
core, Uh = tucker_hooi(dtensor(X1), [3, 4, 2], init='random')
normres = np.sqrt(np.linalg.norm(X1) ** 2 - np.linalg.norm(core) ** 2)
J['hooi_cfg1_rand3'] = {'core_norm': float(np.linalg.norm(core)), 'fit': float(1 - normres / np.linalg.norm(X1))}
Xh = rng.rand(20, 18, 16)
put('hooi_b_X', Xh)
core, Uh = tucker_hooi(dtensor(Xh), [4, 5, 3], init='nvecs', maxIter=6, conv=0)
put('hooi_b_core', np.asarray(core))
for m in range(3):
    put('hooi_b_U%d' % m, Uh[m])
J['hooi_b'] = {'core_norm': float(np.linalg.norm(core))}

# ---------------------------------------------------------------- RESCAL-ALS (rescal.py:41-265)
from scipy.sparse import csr_matrix  # noqa: E402
from sktensor import rescal as ref_rescal  # noqa: E402

rs = np.random.RandomState(77)
J['rescal_cases'] = {}
for name, (n_, k_, r_, la_, lr_, dens_) in {'a': (40, 3, 4, 0.1, 0.2, 0.10), 'b': (60, 2, 5, 0.0, 0.0, 0.08),
                                            'c': (25, 4, 3, 1.0, 0.0, 0.15)}.items():
    Xs = [(rs.rand(n_, n_) < dens_) * rs.rand(n_, n_) for _ in range(k_)]
    for i_, Xi_ in enumerate(Xs):
        put('rescal_%s_X%d' % (name, i_), Xi_)
    np.random.seed(5)
    A_, R_, f_, itr_, _ = ref_rescal.als([csr_matrix(Xi_) for Xi_ in Xs], r_, init='random', lambda_A=la_, lambda_R=lr_)
    put('rescal_%s_A' % name, A_)
    put('rescal_%s_R' % name, np.array(R_))
    J['rescal_cases'][name] = {'n': n_, 'k': k_, 'rank': r_, 'lambda_A': la_, 'lambda_R': lr_, 'seed': 5, 'itr': int(itr_),
                               'f': float(f_)}
    # default init ('nvecs' through ARPACK: column signs are arbitrary) -> record the sign-invariant reconstruction
    A_, R_, f_, itr_, _ = ref_rescal.als([csr_matrix(Xi_) for Xi_ in Xs], r_, lambda_A=la_, lambda_R=lr_)
    put('rescal_%s_nvecs_recon' % name, np.array([A_.dot(Rk_).dot(A_.T) for Rk_ in R_]))
    J['rescal_cases'][name]['itr_nvecs'] = int(itr_)

J['versions'] = {'numpy': np.__version__, 'scipy': __import__('scipy').__version__,
                 'python': sys.version.split()[0]}

if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=os.path.join(os.path.dirname(HERE), 'tests', 'golden'))
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    np.savez_compressed(os.path.join(args.out, 'golden.npz'), **A)
    with open(os.path.join(args.out, 'golden.json'), 'w') as f:
        json.dump(J, f, indent=1, sort_keys=True)
    print('wrote %d arrays, %d json keys to %s' % (len(A), len(J), args.out))

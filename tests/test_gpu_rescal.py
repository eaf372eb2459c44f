"""RESCAL-ALS on the device (sktensor/rescal.py) against the reference's golden vectors and the oracle restatement
(reference: rescal.py:41-265)."""
import numpy as np
import pytest
from scipy.sparse import csr_matrix

from conftest import relerr
from oracle import sktensor_oracle as orc

pytestmark = pytest.mark.gpu


def test_rescal_random_init_matches_reference(golden):
    from sktensor import rescal
    for name, c in golden.j['rescal_cases'].items():
        X = [csr_matrix(golden['rescal_%s_X%d' % (name, i)]) for i in range(c['k'])]
        np.random.seed(c['seed'])
        A, R, f, itr, ex = rescal.als(X, c['rank'], init='random', lambda_A=c['lambda_A'], lambda_R=c['lambda_R'])
        assert itr == c['itr'] and f == 0 and len(ex) == itr, name
        assert relerr(A, golden['rescal_%s_A' % name]) < 1e-7
        assert relerr(np.array(R), golden['rescal_%s_R' % name]) < 1e-6


def test_rescal_default_init_matches_reference(golden):
    """init='nvecs': eigsh of sum_k X_k + X_k^T in the reference; eigenvector signs are arbitrary there, the
    reconstruction A R_k A^T and the iteration count are not."""
    from sktensor import rescal
    for name, c in golden.j['rescal_cases'].items():
        X = [csr_matrix(golden['rescal_%s_X%d' % (name, i)]) for i in range(c['k'])]
        A, R, f, itr, _ = rescal.als(X, c['rank'], lambda_A=c['lambda_A'], lambda_R=c['lambda_R'])
        assert itr == c['itr_nvecs'], name
        recon = np.array([A.dot(Rk).dot(A.T) for Rk in R])
        assert relerr(recon, golden['rescal_%s_nvecs_recon' % name]) < 1e-6


def test_rescal_midsize_vs_oracle_with_options():
    from sktensor import rescal
    rng = np.random.RandomState(4)
    n, k, r = 3000, 4, 12
    X = []
    for _ in range(k):
        rows, cols = rng.randint(0, n, 40000), rng.randint(0, n, 40000)
        X.append(csr_matrix((np.ones(40000), (rows, cols)), shape=(n, n)))
    np.random.seed(9)
    A, R, f, itr, _ = rescal.als(X, r, init='random', lambda_A=10.0, lambda_R=10.0, maxIter=15, conv=1e-9)
    np.random.seed(9)
    Ao, Ro, fo, io = orc.rescal_als(X, r, np.random.rand(n, r), lambda_A=10.0, lambda_R=10.0, maxIter=15, conv=1e-9)
    assert itr == io
    assert relerr(A, Ao) < 1e-7 and relerr(np.array(R), np.array(Ro)) < 1e-6
    # orthogonalised variant: A has orthonormal columns, objective of rescal.py:274-278
    np.random.seed(9)
    A2, R2, _, _, _ = rescal.als(X, r, init='random', orthogonalize=True, maxIter=5)
    assert relerr(A2.T.dot(A2), np.eye(r)) < 1e-10
    # attributes (rescal.py:224-226,249-262) keep the call running and change the factors
    P = [csr_matrix((np.ones(5000), (rng.randint(0, n, 5000), rng.randint(0, 50, 5000))), shape=(n, 50))]
    np.random.seed(9)
    A3, R3, _, _, _ = rescal.als(X, r, init='random', attr=P, lambda_A=10.0, lambda_R=10.0, lambda_V=1.0, maxIter=5)
    assert np.isfinite(A3).all() and relerr(A3, A) > 1e-6
    with pytest.raises(ValueError):
        rescal.als(X, r, bogus=1)

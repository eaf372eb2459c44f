"""cp_als through the one-call loop (skt_cp_als_run, csrc/cpals_run.cu): small dense tensors, the whole of
cp.py:132-171 behind one C call.  Same golden vectors and oracle as the engine tests; the two routes must agree."""
import os

import numpy as np
import pytest

from conftest import relerr
from oracle import sktensor_oracle as orc

pytestmark = [pytest.mark.gpu, pytest.mark.runloop]


def test_runloop_is_taken_and_matches_golden(golden):
    from sktensor import _lib, cp_als, dtensor
    X = golden['cfg1_X']
    lib = _lib.load()
    np.random.seed(42)
    before = lib.skt_launch_count()
    P, fit, itr, ex = cp_als(dtensor(X), 3, init='random')
    launches = lib.skt_launch_count() - before
    g = golden.j['cp_cfg1_s42']
    assert itr == g['itr'] and abs(float(fit) - g['fit']) < 1e-9 and len(ex) == itr + 1
    assert launches <= (itr + 1) * 8 + 12                       # ~2 launches per mode and iteration, not the engine's ~14
    for m in range(3):
        assert relerr(np.asarray(P.U[m]), golden['cp_cfg1_s42_U%d' % m]) < 1e-7
    assert relerr(np.asarray(P.lmbda), golden['cp_cfg1_s42_lmbda']) < 1e-8


@pytest.mark.parametrize('k', range(10))
def test_runloop_equals_engine_and_oracle(k, monkeypatch):
    from sktensor import cp_als, dtensor
    rng = np.random.RandomState(900 + k)
    N = int(rng.choice([2, 3, 3, 4, 5]))
    shape = tuple(int(rng.randint(2, 40 if N <= 3 else 12)) for _ in range(N))
    R = int(rng.choice([1, 2, 3, 5, 8, 16, 33, 64]))
    X = rng.rand(*shape)
    kw = dict(init='random', max_iter=int(rng.randint(1, 12)), conv=0 if k % 2 else 1e-4)
    np.random.seed(k)
    P1, f1, i1, _ = cp_als(dtensor(X), R, **kw)
    monkeypatch.setenv('SKT_NO_RUN_LOOP', '1')
    np.random.seed(k)
    P2, f2, i2, _ = cp_als(dtensor(X), R, **kw)
    np.random.seed(k)
    Uo, lo, fo, io, _ = orc.cp_als(X, R, **kw)
    assert i1 == i2 == io
    assert abs(float(f1) - float(f2)) < 1e-9 and abs(float(f1) - float(np.ravel(fo)[0])) < 1e-6


def test_runloop_contract_details(golden):
    from sktensor import cp_als, dtensor
    X = golden['cfg1_X']
    # fit_method=None: fit is the iteration index and max_iter sweeps run (cp.py:160-161)
    np.random.seed(42)
    _, fit, itr, ex = cp_als(dtensor(X), 3, init='random', fit_method=None, max_iter=4)
    assert (fit, itr, len(ex)) == (3, 3, 4)
    # a list init is updated in place and aliased by the result (cp.py:154,195-196)
    rng = np.random.RandomState(0)
    U0 = [None] + [rng.rand(s, 3) for s in X.shape[1:]]
    P, _, _, _ = cp_als(dtensor(X), 3, init=U0, max_iter=3, conv=0)
    assert all(P.U[m] is U0[m] for m in range(3)) and U0[0] is not None and U0[0].shape == (10, 3)
    with pytest.raises(ValueError):
        cp_als(dtensor(X), 3, init=[None, np.ones((11, 2)), np.ones((8, 3))])
    # the default init ('nvecs') and golden KAT
    P, fit, itr, _ = cp_als(dtensor(X), 3)
    g = golden.j['cp_cfg1_nvecs']
    assert itr == g['itr'] and abs(float(fit) - g['fit']) < 1e-8
    # non-finite data: scipy's pinv raises ValueError in the reference (cp.py:147)
    Xn = X.copy()
    Xn[1, 2, 3] = np.nan
    np.random.seed(1)
    with pytest.raises(ValueError):
        cp_als(dtensor(Xn), 3, init='random', max_iter=3)

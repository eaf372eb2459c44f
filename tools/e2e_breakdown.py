"""Host wall-clock breakdown of the public cp_als call at BASELINE config 2 (development aid)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'scikit-tensor_b200'))
import torch  # noqa: E402
from sktensor import _parallel, cp_als, dtensor  # noqa: E402

X = np.random.RandomState(0).rand(512, 512, 512)
for rep in range(4):
    if rep == 3:
        os.environ['SKT_NO_CLUSTER_UPDATE'] = '1'
    np.random.seed(1)
    U = [None] + [np.random.rand(512, 32) for _ in range(2)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng = _parallel.CpAlsEngine(dtensor(X), 32, U)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for itr in range(10):
        eng.sweep(first_iter=(itr == 0))
        eng.begin_read()
        if itr < 9:
            eng.prefetch_next_sweep()
        eng.read_fit()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    lam = eng.download(U)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    del eng
    np.random.seed(1)
    t4 = time.perf_counter()
    P, fit, itr, ex = cp_als(dtensor(X), 32, init='random', max_iter=10, conv=0)
    torch.cuda.synchronize()
    t5 = time.perf_counter()
    print('rep %d (cluster %s): engine init %.1f ms | 10 sweeps %.1f ms | download %.1f ms | whole cp_als call %.1f ms (exectimes ms: %s)' % (
        rep, os.environ.get('SKT_NO_CLUSTER_UPDATE', '0') != '1', (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t5 - t4) * 1e3,
        ' '.join('%.2f' % (e * 1e3) for e in ex)))

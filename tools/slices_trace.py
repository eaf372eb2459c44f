"""Per-phase trace of one ALS sweep on 512^3 rank 32 with and without the sliced passes (SKT_SWEEP_TRACE output of bench.timed_sweeps)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'scikit-tensor_b200'))
sys.path.insert(0, ROOT)
from sktensor import _device as dev, _parallel, dtensor   # noqa: E402
import bench                                               # noqa: E402

t = dev.torch()
os.environ['SKT_SWEEP_TRACE'] = '1'
side, R = 512, 32
g = t.Generator(device='cuda').manual_seed(1)
Xd = t.rand((side,) * 3, dtype=t.float64, device='cuda', generator=g)
for env in ('', '6'):
    if env:
        os.environ['SKT_DENSE_I8'] = env
    else:
        os.environ.pop('SKT_DENSE_I8', None)
    X = dtensor(np.broadcast_to(np.zeros(1), (side,) * 3))
    X._dev = Xd
    np.random.seed(1)
    U = [None] + [np.random.rand(side, R) for _ in range(2)]
    eng = _parallel.CpAlsEngine(X, R, U)
    ms, fit = bench.timed_sweeps(t, None, 1, eng, 20, 5)
    print('SKT_DENSE_I8=%s: %.4f ms per iteration, fit %.12f' % (env or '-', ms, fit))
    eng.release()

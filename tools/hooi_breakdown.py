"""Device-time breakdown of one HOOI mode update at BASELINE config 4 (384^3 -> 32^3); development aid."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'scikit-tensor_b200'))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from sktensor import _device as dev, _kernels as K  # noqa: E402
from tools.quick_perf import timeit  # noqa: E402

shape = (384, 384, 384)
r = 32
X = torch.rand(*shape, dtype=torch.float64, device='cuda')
U = [torch.rand(s, r, dtype=torch.float64, device='cuda') for s in shape]
for n in range(3):
    others = [m for m in range(3) if m != n]
    Y1, s1 = K.ttm_dense(X, shape, U[others[0]], others[0], True)
    Y2, s2 = K.ttm_dense(Y1, s1, U[others[1]], others[1], True)
    G = K.unfold_gram(Y2, s2, n)
    print('mode %d: ttm1 %.3f ms  ttm2 %.3f ms  gram %.3f ms  eigh %.3f ms  (eigh+D2H+flipsign+H2D host path %.3f ms)' % (
        n, timeit(lambda: K.ttm_dense(X, shape, U[others[0]], others[0], True))[0],
        timeit(lambda: K.ttm_dense(Y1, s1, U[others[1]], others[1], True))[0],
        timeit(lambda: K.unfold_gram(Y2, s2, n))[0],
        timeit(lambda: torch.linalg.eigh(G))[0],
        timeit(lambda: dev.to_device(np.array(dev.to_host(torch.linalg.eigh(G)[1][:, -r:].flip(1).contiguous()))))[0]))

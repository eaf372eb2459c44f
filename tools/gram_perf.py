import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'scikit-tensor_b200')); sys.path.insert(0, ROOT)
import torch
from sktensor import _device as dev, _kernels as K, cp_als, dtensor
from tools.quick_perf import timeit
shape = (512, 512, 512)
X = torch.rand(*shape, dtype=torch.float64, device='cuda')
for n in range(3):
    med, _ = timeit(lambda: K.unfold_gram(X, shape, n), reps=3, warm=1)
    print('unfold_gram 512^3 mode %d: %.2f ms  (%.1f TFLOP/s)' % (n, med, 2 * 512 * 512**3 / med / 1e9))
A = X.view(512, -1)
med, _ = timeit(lambda: A @ A.T, reps=3, warm=1)
print('cuBLAS A A^T mode 0: %.2f ms (%.1f TFLOP/s)' % (med, 2 * 512 * 512**3 / med / 1e9))
G = K.unfold_gram(X, shape, 0)
med, _ = timeit(lambda: torch.linalg.eigh(G), reps=3, warm=1)
print('eigh 512: %.2f ms' % med)
Xh = np.random.RandomState(0).rand(*shape)
for rep in range(2):
    t0 = time.perf_counter()
    P, fit, itr, _ = cp_als(dtensor(Xh), 32, max_iter=10, conv=0)
    torch.cuda.synchronize()
    print('cp_als default init (nvecs), 10 iterations, whole call: %.1f ms' % ((time.perf_counter() - t0) * 1e3))

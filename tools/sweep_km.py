import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'scikit-tensor_b200'))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sktensor import _device as dev, _kernels as K
from tools.quick_perf import timeit
shape = tuple(int(x) for x in sys.argv[1].split('x')); R = int(sys.argv[2]); mode = int(sys.argv[3])
X = torch.rand(*shape, dtype=torch.float64, device='cuda')
U = [torch.rand(s, R, dtype=torch.float64, device='cuda') for s in shape]
out = dev.empty((shape[mode], R))
med, best = timeit(lambda: K.mttkrp_dense(X, shape, U, R, mode, out=out))
print('SKT_KM_LT=%s shape %s R=%d mode %d: median %.3f ms' % (os.environ.get('SKT_KM_LT'), shape, R, mode, med))

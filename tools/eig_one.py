"""One skt_symeig_top call (n = 384, k = 32 on the Gram of uniform data) -- the ncu target for the cluster eigensolver."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'scikit-tensor_b200'))
import torch  # noqa: E402
from sktensor import _device as dev, _kernels as K  # noqa: E402

n, cols, k = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (384, 1024, 32)))
Y = np.random.RandomState(0).rand(n, cols)
G = dev.to_device(Y.dot(Y.T))
for _ in range(2):
    K.symeig_top(G, k)
torch.cuda.synchronize()

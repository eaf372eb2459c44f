"""Small driver for ncu captures of the sparse MTTKRP kernel."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'scikit-tensor_b200'))
import torch  # noqa: E402
from sktensor import _device as dev, _kernels as K  # noqa: E402

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
nnz = int(sys.argv[2]) if len(sys.argv) > 2 else 20000000
R = int(sys.argv[3]) if len(sys.argv) > 3 else 16
shape = (10000000, 1000000, 100000)
rng = np.random.default_rng(0)
subs = [np.minimum((I * rng.random(nnz) ** 3.0).astype(np.int64), I - 1) for I in shape]
vals = rng.random(nnz)
coo = K.CooHandle(subs, vals, shape)
U = [torch.rand(s, R, dtype=torch.float64, device='cuda') for s in shape]
out = dev.empty((shape[mode], R))
for _ in range(3):
    K.mttkrp_coo(coo, U, R, mode, out=out)
torch.cuda.synchronize()
print('done', float(out.sum()))

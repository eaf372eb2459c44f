"""The three sparse MTTKRPs of BASELINE configs[2] (200 M nnz by default), one launch each after a warm-up -- ncu target."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'scikit-tensor_b200'))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from sktensor import _device as dev, _kernels as K  # noqa: E402

nnz = int(sys.argv[1]) if len(sys.argv) > 1 else 200000000
subs, vals = bench.device_powerlaw_coo(torch, bench.SPARSE_SHAPE, nnz)
coo = K.CooHandle.from_device(subs, vals, bench.SPARSE_SHAPE)
del subs, vals
g = torch.Generator(device='cuda')
g.manual_seed(1)
U = [torch.rand(s, 16, dtype=torch.float64, device='cuda', generator=g) for s in bench.SPARSE_SHAPE]
for rep in range(2):
    for n in range(3):
        K.mttkrp_coo(coo, U, 16, n)
torch.cuda.synchronize()

"""Small driver for ncu captures of the dense MTTKRP kernel (one mode, a few launches)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'scikit-tensor_b200'))
import torch  # noqa: E402
from sktensor import _device as dev, _kernels as K  # noqa: E402

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
R = int(sys.argv[2]) if len(sys.argv) > 2 else 32
shape = tuple(int(s) for s in sys.argv[3].split('x')) if len(sys.argv) > 3 else (512, 512, 512)
X = torch.rand(*shape, dtype=torch.float64, device='cuda')
U = [torch.rand(s, R, dtype=torch.float64, device='cuda') for s in shape]
out = dev.empty((shape[mode], R))
for _ in range(4):
    K.mttkrp_dense(X, shape, U, R, mode, out=out)
torch.cuda.synchronize()
print('done', float(out.sum()))

"""ncu target: a few sliced contractions (both operand majors) of a 512^3 rank-32 tensor."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'scikit-tensor_b200'))
from sktensor import _device as dev, _kernels as K   # noqa: E402

t = dev.torch()
side, R = 512, 32
g = t.Generator(device='cuda').manual_seed(1)
X = t.rand((side, side, side), dtype=t.float64, device='cuda', generator=g)
U = t.rand((side, R), dtype=t.float64, device='cuda', generator=g)
P = K.PlanesHandle(X, (side, side, side), int(sys.argv[1]) if len(sys.argv) > 1 else 6)
T = dev.empty((side * side, R))
for _ in range(3):
    P.ttm(2, U, R, out=T)
    P.ttm(0, U, R, out=T)
t.cuda.synchronize()
P.close()

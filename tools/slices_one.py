"""ncu target: a few sliced contractions.  Default: both operand majors of a 512^3 rank-32 tensor (resident factor planes);
`cfg5`: the two tree passes of 160^4 rank 64 (streamed Khatri-Rao planes, 64 columns per slice)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'scikit-tensor_b200'))
from sktensor import _device as dev, _kernels as K   # noqa: E402

t = dev.torch()
g = t.Generator(device='cuda').manual_seed(1)
if len(sys.argv) > 1 and sys.argv[1] == 'cfg5':
    side, R = 160, 64
    shape = (side,) * 4
    X = t.rand(shape, dtype=t.float64, device='cuda', generator=g)
    U = [t.rand((side, R), dtype=t.float64, device='cuda', generator=g) for _ in range(4)]
    P = K.PlanesHandle(X, shape, 6)
    TA, TB = dev.empty((side * side, R)), dev.empty((side * side, R))
    for _ in range(3):
        P.tree_pass(2, 0, U, R, out=TA)
        P.tree_pass(2, 1, U, R, out=TB)
else:
    side, R = 512, 32
    X = t.rand((side, side, side), dtype=t.float64, device='cuda', generator=g)
    U = t.rand((side, R), dtype=t.float64, device='cuda', generator=g)
    P = K.PlanesHandle(X, (side, side, side), int(sys.argv[1]) if len(sys.argv) > 1 else 6)
    T = dev.empty((side * side, R))
    for _ in range(3):
        P.ttm(2, U, R, out=T)
        P.ttm(0, U, R, out=T)
t.cuda.synchronize()
P.close()

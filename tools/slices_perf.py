"""Timing of the sliced INT8 contraction (skt_planes_ttm) against the FP64 DMMA pass on the two 2-way views of 512^3."""
import os
import sys


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'scikit-tensor_b200'))
from sktensor import _device as dev, _kernels as K   # noqa: E402

t = dev.torch()
side = int(sys.argv[1]) if len(sys.argv) > 1 else 512
R = int(sys.argv[2]) if len(sys.argv) > 2 else 32
shape = (side, side, side)
g = t.Generator(device='cuda').manual_seed(1)
X = t.rand(shape, dtype=t.float64, device='cuda', generator=g)
U = [t.rand((side, R), dtype=t.float64, device='cuda', generator=g) for _ in range(3)]


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    t.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


TA = dev.empty((side * side, R))
TB = dev.empty((side * side, R))
M2 = dev.empty((side, R))
ms_a = timed(lambda: K.mttkrp_dense(X, (side * side, side), [None, U[2]], R, 0, out=TA))
ms_b = timed(lambda: K.mttkrp_dense(X, shape, [U[0], U[1], None], R, 2, out=M2))
print('DMMA: pass A (last-mode contraction, row store) %.4f ms, mode-2 MTTKRP %.4f ms' % (ms_a, ms_b))
refA = TA.clone()
K.mttkrp_dense(X, (side, side * side), [U[0], None], R, 1, out=TB)
refB = TB.clone()
for S in (6, 5):
    e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    e0.record()
    P = K.PlanesHandle(X, shape, S)
    e1.record()
    t.cuda.synchronize()
    print('S=%d: planes built in %.3f ms (%.2f GB)' % (S, e0.elapsed_time(e1), P.device_bytes() / 1e9))
    a = timed(lambda: P.ttm(2, U[2], R, out=TA))
    b = timed(lambda: P.ttm(0, U[0], R, out=TB))
    ea = float((TA - refA).abs().max() / refA.abs().max())
    eb = float((TB - refB).abs().max() / refB.abs().max())
    c = timed(lambda: K.kr_reduce(TB, (side, side), [U[1], None], R, 1, out=M2))
    print('  last-mode %.4f ms (%.0f GB/s of planes, err %.2e)   first-mode %.4f ms (err %.2e)   + kr_reduce %.4f ms'
          % (a, S * side ** 3 / a * 1e-6, ea, b, eb, c))
    # the general pass on the same tensor: [512^2 x 512] against U_2 (resident planes) and the transposed product against KR(U_0, U_1)
    TA2 = P.tree_pass(2, 0, [None, None, U[2]], R)
    a2 = timed(lambda: P.tree_pass(2, 0, [None, None, U[2]], R, out=TA2))
    M2b = P.tree_pass(2, 1, [U[0], U[1], None], R)
    b2 = timed(lambda: P.tree_pass(2, 1, [U[0], U[1], None], R, out=M2b))
    K.mttkrp_dense(X, shape, [U[0], U[1], None], R, 2, out=M2)
    print('  general pass: group 0 %.4f ms (err %.2e)   group 1 (K = 512^2, split-K) %.4f ms (err %.2e)'
          % (a2, float((TA2 - refA).abs().max() / refA.abs().max()), b2, float((M2b - M2).abs().max() / M2.abs().max())))
    P.close()

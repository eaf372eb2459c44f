"""Quick device-side timings (CUDA events) of the hot kernels; development aid, not the bench."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'scikit-tensor_b200'))
import torch  # noqa: E402
from sktensor import _device as dev, _kernels as K  # noqa: E402


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def main():
    print('fp64 peaks (dmma, dfma) TFLOP/s:', K.probe_fp64_peaks())
    a = torch.rand(8192, 8192, dtype=torch.float64, device='cuda')
    b = torch.rand(8192, 8192, dtype=torch.float64, device='cuda')
    med, best = timeit(lambda: torch.matmul(a, b), reps=5, warm=2)
    print('cuBLAS dgemm 8192^3: %.2f ms -> %.1f TFLOP/s' % (best, 2 * 8192 ** 3 / best / 1e9))
    del a, b
    which = sys.argv[1:] or ['cfg2', 'cfg5', 'sparse']
    if 'cfg2' in which:
        for shape, R in (((512, 512, 512), 32), ((512, 512, 512), 16), ((512, 512, 512), 64)):
            X = torch.rand(*shape, dtype=torch.float64, device='cuda')
            U = [torch.rand(s, R, dtype=torch.float64, device='cuda') for s in shape]
            nbytes = 8 * (X.numel() + R * sum(shape))
            for n in range(len(shape)):
                out = dev.empty((shape[n], R))
                med, best = timeit(lambda: K.mttkrp_dense(X, shape, U, R, n, out=out))
                print('dense %s R=%d mode %d: median %.3f ms best %.3f ms  %.0f GB/s  %.1f TFLOP/s' %
                      (shape, R, n, med, best, nbytes / med / 1e6, 2.0 * R * X.numel() / med / 1e9))
            del X
    if 'cfg5' in which:
        shape, R = (160, 160, 160, 160), 64
        X = torch.rand(*shape, dtype=torch.float64, device='cuda')
        U = [torch.rand(s, R, dtype=torch.float64, device='cuda') for s in shape]
        nbytes = 8 * (X.numel() + R * sum(shape))
        for n in range(4):
            out = dev.empty((shape[n], R))
            med, best = timeit(lambda: K.mttkrp_dense(X, shape, U, R, n, out=out), reps=5)
            print('dense %s R=%d mode %d: median %.3f ms best %.3f ms  %.0f GB/s  %.1f TFLOP/s' %
                  (shape, R, n, med, best, nbytes / med / 1e6, 2.0 * R * X.numel() / med / 1e9))
        del X
    if 'sparse' in which:
        shape = (10000000, 1000000, 100000)
        nnz = int(os.environ.get('NNZ', 20000000))
        R = 16
        rng = np.random.default_rng(0)
        subs = [np.minimum((I * rng.random(nnz) ** 3.0).astype(np.int64), I - 1) for I in shape]
        vals = rng.random(nnz)
        t0 = time.perf_counter()
        coo = K.CooHandle(subs, vals, shape)
        torch.cuda.synchronize()
        print('coo build (%d nnz): %.2f s, %.2f GB on device' % (nnz, time.perf_counter() - t0, coo.device_bytes() / 1e9))
        U = [torch.rand(s, R, dtype=torch.float64, device='cuda') for s in shape]
        for n in range(3):
            out = dev.empty((shape[n], R))
            med, best = timeit(lambda: K.mttkrp_coo(coo, U, R, n, out=out), reps=5)
            nbytes = nnz * (4 * 3 + 8) + 8 * R * sum(shape)
            print('sparse nnz=%d R=%d mode %d: median %.3f ms best %.3f ms  %.0f GB/s (algorithmic)  %.1f Mnnz/s' %
                  (nnz, R, n, med, best, nbytes / med / 1e6, nnz / med / 1e3))
    if 'tree' in which:
        for shape, R in (((512, 512, 512), 32), ((160, 160, 160, 160), 64)):
            N = len(shape)
            X = torch.rand(*shape, dtype=torch.float64, device='cuda')
            U = [torch.rand(s, R, dtype=torch.float64, device='cuda') for s in shape]
            s_ = 2
            nA, nB = int(np.prod(shape[:s_])), int(np.prod(shape[s_:]))
            shapeA, shapeB = (nA,) + tuple(shape[s_:]), tuple(shape[:s_]) + (nB,)
            TA, TB = dev.empty((nA, R)), dev.empty((nB, R))
            med, best = timeit(lambda: K.mttkrp_dense(X, shapeA, [None] + U[s_:], R, 0, out=TA), reps=5)
            print('tree %s R=%d pass A (%s mode 0): median %.3f ms  %.1f TFLOP/s' % (shape, R, shapeA, med, 2.0 * R * X.numel() / med / 1e9))
            med, best = timeit(lambda: K.mttkrp_dense(X, shapeB, U[:s_] + [None], R, s_, out=TB), reps=5)
            print('tree %s R=%d pass B (%s mode %d): median %.3f ms  %.1f TFLOP/s' % (shape, R, shapeB, s_, med, 2.0 * R * X.numel() / med / 1e9))
            for n in range(s_):
                out = dev.empty((shape[n], R))
                med, best = timeit(lambda: K.kr_reduce(TA, shape[:s_], U[:s_], R, n, out=out))
                print('  kr_reduce group A mode %d: median %.4f ms (%.0f GB/s over T)' % (n, med, TA.numel() * 8 / med / 1e6))
            if s_ < N - 1:
                for n in range(s_, N):
                    out = dev.empty((shape[n], R))
                    med, best = timeit(lambda: K.kr_reduce(TB, shape[s_:], U[s_:], R, n - s_, out=out))
                    print('  kr_reduce group B mode %d: median %.4f ms (%.0f GB/s over T)' % (n, med, TB.numel() * 8 / med / 1e6))
            del X, TA, TB
    if 'small' in which:
        for rows, R, N in ((512, 32, 3), (160, 64, 4)):
            rng = np.random.RandomState(0)
            G = dev.to_device(np.stack([(lambda f: f.T.dot(f))(rng.rand(rows, R)) for _ in range(N)]))
            M = dev.to_device(rng.rand(rows, R))
            U, lam, info = dev.empty((rows, R)), dev.empty((R,)), dev.zeros((1,), np.int32)
            P, colstat = dev.empty((R, R)), dev.empty((R,))
            ip, fit, nx2 = dev.zeros((1,)), dev.zeros((2,)), dev.to_device(np.array([1e6]))
            G0 = G.clone()
            def fused():
                K.cp_update_fused(M, G, 0, False, U, lam, info, ip=ip, normX2=nx2, fit=fit)
            print('rows=%d R=%d: fused update            %.2f us' % (rows, R, 1e3 * timeit(fused, reps=20)[0]))
            G.copy_(G0)
            Pd, _ = K.hadamard_pinv(G, 0)
            def clus():
                K.cp_update_cluster(M, Pd, G, 0, False, U, lam, ip=ip, normX2=nx2, fit=fit)
            print('             cluster update (no pinv) %.2f us' % (1e3 * timeit(clus, reps=20)[0]))
            G.copy_(G0)
            print('             hadamard_pinv           %.2f us' % (1e3 * timeit(lambda: K.hadamard_pinv(G, 0, P=P, info=info), reps=20)[0]))
            print('             solve_stats             %.2f us' % (1e3 * timeit(lambda: K.solve_stats(M, P, False, Unew=U, colstat=colstat), reps=20)[0]))
            print('             normalise_gram          %.2f us' % (1e3 * timeit(lambda: K.normalise_gram(U, colstat, False, lam=lam, G=G[0], Mip=M, ip=ip), reps=20)[0]))
            print('             cp_fit                  %.2f us' % (1e3 * timeit(lambda: K.cp_fit(G, lam, ip, nx2, out=fit), reps=20)[0]))
            e = dev.empty((1,))
            print('             (empty torch op)        %.2f us' % (1e3 * timeit(lambda: e.zero_(), reps=20)[0]))
    if 'upload' in which:
        h = np.random.RandomState(0).rand(512, 512, 512)
        for rep in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            x = dev.to_device(h)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            y = torch.from_numpy(h).to('cuda')
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            print('upload 1.07 GB: skt_upload %.1f ms (%.1f GB/s)   torch pageable %.1f ms (%.1f GB/s)  equal=%s' %
                  ((t1 - t0) * 1e3, h.nbytes / (t1 - t0) / 1e9, (t2 - t1) * 1e3, h.nbytes / (t2 - t1) / 1e9,
                   bool(torch.equal(x, y))))
            del x, y
    if 'als4' in which:
        from sktensor import _parallel, dtensor
        X = np.random.RandomState(0).rand(160, 160, 160, 160)
        np.random.seed(1)
        U = [None] + [np.random.rand(160, 64) for _ in range(3)]
        for flag in ('1', '0'):
            os.environ['SKT_NO_DIMTREE'] = flag
            eng = _parallel.CpAlsEngine(dtensor(X), 64, list(U))
            eng.sweep(True)
            f0 = eng.read_fit()
            med, best = timeit(lambda: (eng.sweep(False), eng.read_fit()), reps=6, warm=2)
            print('cp_als 160^4 R=64 iteration (dimtree %s): median %.3f ms best %.3f ms -> %.1f it/s (fit %.12f)' %
                  ('off' if flag == '1' else 'on', med, best, 1e3 / med, eng.read_fit()))
            del eng
        os.environ['SKT_NO_DIMTREE'] = '0'
    if 'hooi' in which:
        from sktensor import dtensor, tucker_hooi
        X = np.random.RandomState(0).rand(384, 384, 384)
        for rep in range(2):
            np.random.seed(0)
            t0 = time.perf_counter()
            core, Uh = tucker_hooi(dtensor(X), [32, 32, 32], init='random', maxIter=5, conv=0)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            print('tucker_hooi 384^3 -> 32^3, 5 iterations, whole call: %.1f ms -> %.2f it/s' % (dt * 1e3, 5 / dt))
    if 'als' in which:
        os.environ['SKT_NO_DIMTREE'] = os.environ.get('SKT_NO_DIMTREE', '0')
        from sktensor import _parallel, dtensor
        X = np.random.RandomState(0).rand(512, 512, 512)
        np.random.seed(1)
        U = [None] + [np.random.rand(512, 32) for _ in range(2)]
        eng = _parallel.CpAlsEngine(dtensor(X), 32, U)
        eng.sweep(True)
        eng.read_fit()
        med, best = timeit(lambda: (eng.sweep(False), eng.read_fit()), reps=10, warm=2)
        print('cp_als 512^3 R=32 iteration: median %.3f ms best %.3f ms -> %.1f it/s' % (med, best, 1e3 / med))


if __name__ == '__main__':
    main()

"""Timing of skt_symeig_top against the library eigensolver, and of the whole tucker_hooi call at config 4."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'scikit-tensor_b200'))
import torch  # noqa: E402
from sktensor import _device as dev, _kernels as K, dtensor, tucker_hooi  # noqa: E402

out = {'symeig': [], 'hooi': {}}
rng = np.random.RandomState(0)
for n, cols, k in [(384, 1024, 32), (512, 4096, 32), (160, 4096, 64), (128, 128, 128), (1024, 2048, 16)]:
    Y = rng.rand(n, cols)
    G = dev.to_device(Y.dot(Y.T))
    for _ in range(3):
        K.symeig_top(G, k)
    torch.cuda.synchronize()
    os.environ['SKT_EIG_PROF'] = '1'
    K.symeig_top(G, k)
    os.environ['SKT_EIG_PROF'] = '0'
    sys.stderr.flush()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); K.symeig_top(G, k); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    tl = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.linalg.eigh(G); e1.record(); torch.cuda.synchronize()
        tl.append(e0.elapsed_time(e1))
    out['symeig'].append({'n': n, 'k': k, 'ms': float(np.median(ts)), 'library_eigh_ms': float(np.median(tl))})
    print(out['symeig'][-1], flush=True)

X = dtensor(np.random.RandomState(0).rand(384, 384, 384))
for init in ('random', 'nvecs'):
    for it in (5, 20):
        np.random.seed(0)
        tucker_hooi(X, [32, 32, 32], init=init, maxIter=2, conv=0)
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            np.random.seed(0)
            tic = time.perf_counter()
            tucker_hooi(X, [32, 32, 32], init=init, maxIter=it, conv=0)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - tic)
        out['hooi']['%s_%d' % (init, it)] = {'seconds': float(np.median(ts)), 'its_per_s': it / float(np.median(ts))}
        print(init, it, out['hooi']['%s_%d' % (init, it)], flush=True)
# steady-state iteration cost: difference between 20 and 5 iterations
for init in ('random', 'nvecs'):
    d = (out['hooi']['%s_20' % init]['seconds'] - out['hooi']['%s_5' % init]['seconds']) / 15.0
    out['hooi']['%s_ms_per_iteration_steady' % init] = d * 1e3
print(json.dumps(out))

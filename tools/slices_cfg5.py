"""cfg5 (dense 160^4, rank 64) with the tree passes on the INT8 tensor cores (SKT_DENSE_I8) against the FP64 DMMA engine."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'scikit-tensor_b200'))
sys.path.insert(0, ROOT)
from sktensor import _device as dev, _parallel, dtensor   # noqa: E402
import bench                                                              # noqa: E402

t = dev.torch()
R = 64


def engine(env):
    if env:
        os.environ['SKT_DENSE_I8'] = env
    else:
        os.environ.pop('SKT_DENSE_I8', None)
    Xt, gshape, smode = bench.device_shard_cfg5(t, _parallel, dtensor, 1, 0)
    np.random.seed(1)
    U = [None] + [np.random.rand(gshape[n], R) for n in range(1, 4)]
    return _parallel.CpAlsEngine(Xt, R, U), Xt


for env in (sys.argv[1].split(',') if len(sys.argv) > 1 else ['', '6', '5']):
    eng, Xt = engine(env)
    ms, fit = bench.timed_sweeps(t, None, 1, eng, 10, 3)
    print('SKT_DENSE_I8=%-2s  %.3f ms per iteration, fit %.12f, sliced=%s general=%s' %
          (env or '-', ms, fit, eng.i8 is not None, getattr(eng, 'i8_general', False)))
    if eng.i8 is not None:
        e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        for grp in (0, 1):
            for _ in range(2):
                eng._tree_pass(grp)
            e0.record()
            for _ in range(5):
                eng._tree_pass(grp)
            e1.record()
            t.cuda.synchronize()
            print('    pass of group %d: %.3f ms' % (grp, e0.elapsed_time(e1) / 5))
    eng.release()
    del eng, Xt
    t.cuda.empty_cache()

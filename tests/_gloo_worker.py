"""Worker for tests/test_distributed_gloo.py (gloo, CPU, NumPy test double) and
tests/test_gpu_multi.py (nccl, one GPU per rank, CUDA kernels): one rank of a world_size-N job."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, 'scikit-tensor_b200'), ROOT, os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)

import torch.distributed as dist  # noqa: E402

from fake_backend import NumpyBackend  # noqa: E402
from sktensor import _parallel, cp_als, dtensor, sptensor  # noqa: E402
from sktensor.sptensor import fromarray  # noqa: E402


def main():
    out_path = sys.argv[1]
    mode = sys.argv[2] if len(sys.argv) > 2 else 'gloo'
    if mode == 'nccl':        # real GPUs: the shipped CUDA backend, one rank per device
        import torch
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        dist.init_process_group('nccl')
    else:                     # CPU: gloo + the NumPy test double
        dist.init_process_group('gloo')
        _parallel.set_backend_factory(NumpyBackend)
    rank = dist.get_rank()
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'golden.npz'))
    res = {}

    def run(key, X, R, seed, **kw):
        np.random.seed(seed)
        P, fit, itr, _ = cp_als(X, R, **kw)
        res[key] = {'fit': float(fit), 'itr': int(itr), 'lmbda': np.asarray(P.lmbda).tolist(),
                    'U': [np.asarray(u).tolist() for u in P.U]}

    X1 = g['cfg1_X']
    run('cp_cfg1_s42', dtensor(X1), 3, 42, init='random')                       # sharded on mode 1 (11)
    run('cp_cfg1_s42_it5', dtensor(X1), 3, 42, init='random', max_iter=5, conv=0)
    run('cp_kat3_s42', fromarray(X1 * (X1 > 0.7)), 3, 42, init='random')        # sparse
    run('cp_d4_s3', dtensor(g['cp4_X']), 4, 3, init='random', max_iter=30)      # 4-way, sharded on the last mode
    subs = tuple(g['cpsp_sub%d' % m] for m in range(3))
    run('cp_sp_s1', sptensor(subs, g['cpsp_vals'], shape=(60, 50, 40)), 5, 1, init='random', max_iter=25)
    # tall replicated factors take the scattered update (reduce-scatter, exchange-form update of the own rows,
    # all-gather): forced here on a small tensor, same golden run
    os.environ['SKT_SCATTER_MIN_ROWS'] = '1'
    os.environ['SKT_NO_CLUSTER_UPDATE'] = '1'
    os.environ['SKT_NO_FUSED_UPDATE'] = '1'
    run('cp_sp_s1_scatter', sptensor(subs, g['cpsp_vals'], shape=(60, 50, 40)), 5, 1, init='random', max_iter=25)
    run('cp_d4_s3_scatter', dtensor(g['cp4_X']), 4, 3, init='random', max_iter=30)
    for k_ in ('SKT_SCATTER_MIN_ROWS', 'SKT_NO_CLUSTER_UPDATE', 'SKT_NO_FUSED_UPDATE'):
        os.environ.pop(k_)
    # lopsided cuts: rank 0 owns every row of the sharded mode, all other ranks own EMPTY slabs (what happens for real
    # when a mode is shorter than the world size, or with power-law non-zero cuts); results must not change
    world = dist.get_world_size()
    offs = np.array([0] + [X1.shape[1]] * world, dtype=np.int64)
    lo, hi = int(offs[rank]), int(offs[rank + 1])
    run('cp_cfg1_s42_lopsided', _parallel.LocalShard(dtensor(np.ascontiguousarray(X1[:, lo:hi, :])), 1, X1.shape, offs),
        3, 42, init='random')
    S3 = fromarray(X1 * (X1 > 0.7))
    keep = (np.asarray(S3.subs[1]) >= lo) & (np.asarray(S3.subs[1]) < hi)
    lsubs = tuple((np.asarray(sm)[keep] - (lo if m == 1 else 0)).astype(np.int64) for m, sm in enumerate(S3.subs))
    lshape = (X1.shape[0], hi - lo, X1.shape[2])
    run('cp_kat3_s42_lopsided', _parallel.LocalShard(sptensor(lsubs, np.asarray(S3.vals)[keep], shape=lshape), 1, X1.shape, offs),
        3, 42, init='random')
    if mode == 'nccl':
        # the same run with the reductions fused into the update kernels over NVLink peer memory (default from 4 ranks up)
        os.environ['SKT_SYMM'] = '1' if world < 4 else '0'
        run('cp_cfg1_s42_othersymm', dtensor(X1), 3, 42, init='random')
        run('cp_d4_s3_othersymm', dtensor(g['cp4_X']), 4, 3, init='random', max_iter=30)
        os.environ.pop('SKT_SYMM')
    np.random.seed(42)
    _, fit, itr, _ = cp_als(dtensor(X1), 3, init='random', fit_method=None, max_iter=4)
    res['nofit'] = {'fit': float(fit), 'itr': int(itr)}
    with open('%s.%d' % (out_path, rank), 'w') as f:
        json.dump(res, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()

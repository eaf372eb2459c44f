"""Worker for tests/test_gpu_multi.py::test_cfg3_shape_ngpu_equals_1gpu: cp_als on a BASELINE configs[2]-shaped
power-law tensor (2e6 nnz), sharded over the ranks of an NCCL job -- or, without torchrun, on one GPU."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, 'scikit-tensor_b200'), ROOT):
    sys.path.insert(0, p)

SHAPE = (10000000, 1000000, 100000)


def main():
    out_path = sys.argv[1]
    world = int(os.environ.get('WORLD_SIZE', '1'))
    import torch
    rank = 0
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        dist.init_process_group('nccl')
        rank = dist.get_rank()
    from sktensor import cp_als, sptensor
    rng = np.random.default_rng(0)
    nnz = 2000000
    subs = tuple(np.minimum((I * rng.random(nnz) ** 3.0).astype(np.int64), I - 1) for I in SHAPE)
    vals = rng.random(nnz)
    np.random.seed(1)
    P, fit, itr, _ = cp_als(sptensor(subs, vals, shape=SHAPE), 16, init='random', max_iter=3, conv=0)
    res = {'fit': float(fit), 'itr': int(itr), 'lmbda': np.asarray(P.lmbda).tolist(),
           'colsum': [np.asarray(u).sum(axis=0).tolist() for u in P.U],
           'head': [np.asarray(u)[:64].tolist() for u in P.U],
           'sq': [float((np.asarray(u) ** 2).sum()) for u in P.U]}
    with open('%s.%d' % (out_path, rank), 'w') as f:
        json.dump(res, f)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

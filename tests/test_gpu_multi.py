"""Sharded CP-ALS on real GPUs (NCCL, one process per device): every rank must reproduce the
reference's golden run.  Skipped on boxes with fewer than 2 GPUs."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, relerr
from test_distributed_gloo import _free_port

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize('world', [2, 4, 8])
def test_sharded_cp_als_nccl(tmp_path, golden, world):
    if _ngpu() < world:
        pytest.skip('needs %d GPUs' % world)
    out = str(tmp_path / 'res.json')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
           '--master-addr', '127.0.0.1', '--master-port', str(_free_port()),
           os.path.join(ROOT, 'tests', '_gloo_worker.py'), out, 'nccl']
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    results = [json.load(open('%s.%d' % (out, k))) for k in range(world)]
    for key in ('cp_cfg1_s42', 'cp_cfg1_s42_it5', 'cp_kat3_s42', 'cp_d4_s3', 'cp_sp_s1', 'cp_sp_s1_scatter', 'cp_d4_s3_scatter', 'cp_cfg1_s42_lopsided',
                'cp_kat3_s42_lopsided', 'cp_cfg1_s42_othersymm', 'cp_d4_s3_othersymm'):
        # lopsided: empty slabs on all ranks but one; othersymm: NVLink peer-memory reductions toggled -- same golden run
        gkey = key.replace('_lopsided', '').replace('_scatter', '').replace('_othersymm', '')
        for res in results:
            got = res[key]
            assert got['itr'] == golden.j[gkey]['itr'], key
            assert abs(got['fit'] - golden.j[gkey]['fit']) < 1e-6, key
            for m, u in enumerate(got['U']):
                assert relerr(np.array(u), golden['%s_U%d' % (gkey, m)]) < 1e-6, (key, m)
        for res in results[1:]:                      # replicas agree bit for bit
            assert res[key]['fit'] == results[0][key]['fit']


@pytest.mark.parametrize('world', [2, 4, 8])
def test_cfg3_shape_ngpu_equals_1gpu(tmp_path, world):
    """BASELINE configs[2] shape (10M x 1M x 100K power law, 2e6 nnz, rank 16): the sharded run over `world` GPUs
    (equal-nnz cuts along mode 0, reduce-scatter of the tall partial MTTKRPs) reproduces the single-GPU run."""
    if _ngpu() < world:
        pytest.skip('needs %d GPUs' % world)
    worker = os.path.join(ROOT, 'tests', '_nccl_cfg3_worker.py')
    one = str(tmp_path / 'one.json')
    r = subprocess.run([sys.executable, worker, one], capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, CUDA_VISIBLE_DEVICES='0'))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    ref = json.load(open(one + '.0'))
    out = str(tmp_path / 'n.json')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
           '--master-addr', '127.0.0.1', '--master-port', str(_free_port()), worker, out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    for k in range(world):
        got = json.load(open('%s.%d' % (out, k)))
        assert got['itr'] == ref['itr']
        assert abs(got['fit'] - ref['fit']) < 1e-9
        assert relerr(got['lmbda'], ref['lmbda']) < 1e-9
        for m in range(3):
            assert relerr(got['colsum'][m], ref['colsum'][m]) < 1e-9
            assert relerr(np.array(got['head'][m]), np.array(ref['head'][m])) < 1e-8
            assert abs(got['sq'][m] - ref['sq'][m]) < 1e-9 * ref['sq'][m]

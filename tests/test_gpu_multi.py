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
    for key in ('cp_cfg1_s42', 'cp_cfg1_s42_it5', 'cp_kat3_s42', 'cp_d4_s3', 'cp_sp_s1', 'cp_cfg1_s42_lopsided',
                'cp_kat3_s42_lopsided', 'cp_cfg1_s42_othersymm', 'cp_d4_s3_othersymm'):
        # lopsided: empty slabs on all ranks but one; othersymm: NVLink peer-memory reductions toggled -- same golden run
        gkey = key.replace('_lopsided', '').replace('_othersymm', '')
        for res in results:
            got = res[key]
            assert got['itr'] == golden.j[gkey]['itr'], key
            assert abs(got['fit'] - golden.j[gkey]['fit']) < 1e-6, key
            for m, u in enumerate(got['U']):
                assert relerr(np.array(u), golden['%s_U%d' % (gkey, m)]) < 1e-6, (key, m)
        for res in results[1:]:                      # replicas agree bit for bit
            assert res[key]['fit'] == results[0][key]['fit']

"""world_size-2 and -3 runs of the sharded CP-ALS choreography on CPU (gloo), checked against
the golden vectors of the unmodified reference.  The kernels are replaced by the NumPy test
double (tests/fake_backend.py); what is under test is sktensor/_parallel.py: the cuts, which
collective follows which kernel, and that every rank ends with the same, correct model."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, relerr


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize('world', [2, 3])
def test_sharded_cp_als_matches_reference(tmp_path, golden, world):
    out = str(tmp_path / 'res.json')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
           '--master-addr', '127.0.0.1', '--master-port', str(_free_port()),
           os.path.join(ROOT, 'tests', '_gloo_worker.py'), out]
    env = dict(os.environ, OMP_NUM_THREADS='1', PYTHONWARNINGS='ignore')
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    results = [json.load(open('%s.%d' % (out, k))) for k in range(world)]
    for key in ('cp_cfg1_s42', 'cp_cfg1_s42_it5', 'cp_kat3_s42', 'cp_d4_s3', 'cp_sp_s1', 'cp_sp_s1_scatter', 'cp_d4_s3_scatter', 'cp_cfg1_s42_lopsided',
                'cp_kat3_s42_lopsided'):
        gkey = key.replace('_lopsided', '').replace('_scatter', '')          # empty slabs on all ranks but one: same golden run
        for res in results:
            got = res[key]
            assert got['itr'] == golden.j[gkey]['itr'], key
            assert abs(got['fit'] - golden.j[gkey]['fit']) < 1e-9, key
            assert relerr(got['lmbda'], golden['%s_lmbda' % gkey]) < 1e-8
            for m, u in enumerate(got['U']):
                assert relerr(np.array(u), golden['%s_U%d' % (gkey, m)]) < 1e-7, (key, m)
        # replicas agree bit for bit
        for res in results[1:]:
            assert res[key]['fit'] == results[0][key]['fit']
            for a, b in zip(res[key]['U'], results[0][key]['U']):
                assert a == b
    assert all(res['nofit'] == {'fit': 3.0, 'itr': 3} for res in results)

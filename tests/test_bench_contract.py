"""bench.py contract checks that need no GPU: the reference arm runs the unmodified reference (baseline/_ref; the oracle port when absent) on the host cores and prints one
JSON line with the keys the driver reads; the GPU arm's helpers compute the algorithmic byte counts of SURVEY 8(d)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS=os.environ.get('OMP_NUM_THREADS', '4'), SKT_BENCH_REF_SIDE='96')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip().startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'cp_als_mttkrp_GBps' and d['unit'] == 'GB/s'
    assert d['higher_is_better'] is True and d['dtype'] == 'f64' and d['value'] > 0 and d['ms_per_step'] > 0
    assert d['cpu_baseline']['kind'] == ('reference' if os.path.isdir(os.path.join(ROOT, 'baseline', '_ref', 'sktensor')) else 'port') and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'GB/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and '512x512x512' in d['config']['workload']


def test_algorithmic_byte_counts():
    sys.path.insert(0, ROOT)
    import bench
    # config 2 -> 1.0741 GB per dense MTTKRP; config 3 -> 5.4208 GB per sparse MTTKRP
    assert abs(bench.alg_bytes_dense((512, 512, 512), 32) - 1.0741e9) < 1e5 * 10
    assert abs(bench.alg_bytes_sparse((10000000, 1000000, 100000), 200000000, 16) - 5.4208e9) < 1e6

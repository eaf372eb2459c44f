"""pytest configuration: the ``gpu`` marker, import paths and the golden-vector fixture."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'scikit-tensor_b200')
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
    config.addinivalue_line('markers', 'runloop: let cp_als take the one-call loop (skt_cp_als_run) for small tensors')


@pytest.fixture(autouse=True)
def _engine_by_default(request, monkeypatch):
    """Small dense cp_als calls normally run behind one C call (skt_cp_als_run).  The bulk of the suite is about the
    sweep engine (dimension tree, cluster / exchange updates, sharding), so it is pinned to the engine here;
    tests marked `runloop` (tests/test_gpu_runloop.py) exercise the one-call loop, against the same references."""
    if request.node.get_closest_marker('runloop') is None:
        monkeypatch.setenv('SKT_NO_RUN_LOOP', '1')
    else:
        monkeypatch.delenv('SKT_NO_RUN_LOOP', raising=False)


class Golden(object):
    """Vectors written by oracle/make_golden.py from the unmodified reference."""

    def __init__(self):
        d = os.path.join(ROOT, 'tests', 'golden')
        self.a = np.load(os.path.join(d, 'golden.npz'))
        with open(os.path.join(d, 'golden.json')) as f:
            self.j = json.load(f)

    def __getitem__(self, k):
        return self.a[k]

    def factors(self, prefix, n):
        return [self.a['%s_U%d' % (prefix, m)] for m in range(n)]


@pytest.fixture(scope='session')
def golden():
    return Golden()


def relerr(a, b):
    """max|a-b| / max|b|  -- the MTTKRP parity measure of SURVEY 8(d)."""
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    den = np.max(np.abs(b)) if b.size else 0.0
    num = np.max(np.abs(a - b)) if b.size else 0.0
    return num / den if den > 0 else num

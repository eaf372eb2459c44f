"""Alias-only compatibility shim that lets the UNMODIFIED reference import on
Python 3.12 / NumPy 2.x / SciPy 1.14+.

Test infrastructure (used by ``oracle/make_golden.py`` and the optional live
cross-check in ``tests/``); never imported by the product.  Nothing here touches
arithmetic: it restores removed aliases the reference evaluates at import time
(``np.float`` at cp.py:37 / core.py:275, ``time.clock`` at cp.py:139, ``xrange``
at core.py:101, ``collections.Sequence`` at pyutils.py:14) and translates the
removed ``eigvals=`` keyword of ``scipy.linalg.eigh`` (core.py:287).
"""
import builtins
import collections
import collections.abc
import sys
import time

import numpy as np
import scipy.linalg

REFERENCE_ROOT = '/root/reference'


def install(reference_root=REFERENCE_ROOT):
    """Install the aliases and put the reference tree first on ``sys.path``."""
    for name, val in (('float', float), ('int', int), ('float_', np.float64),
                      ('Inf', np.inf), ('bool', bool), ('object', object)):
        if not hasattr(np, name):
            setattr(np, name, val)
    if not hasattr(np, 'in1d'):
        np.in1d = np.isin
    if not hasattr(time, 'clock'):
        time.clock = time.process_time
    if not hasattr(builtins, 'xrange'):
        builtins.xrange = range
    if not hasattr(collections, 'Sequence'):
        collections.Sequence = collections.abc.Sequence

    if not getattr(scipy.linalg.eigh, '_skt_shim', False):
        _orig = scipy.linalg.eigh

        def eigh(a, *args, eigvals=None, **kw):
            if eigvals is not None:
                kw['subset_by_index'] = (int(eigvals[0]), int(eigvals[1]))
            return _orig(a, *args, **kw)

        eigh._skt_shim = True
        scipy.linalg.eigh = eigh
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)


def import_reference(reference_root=REFERENCE_ROOT):
    """Return the reference ``sktensor`` module (refuses to alias the product)."""
    install(reference_root)
    if 'sktensor' in sys.modules:
        mod = sys.modules['sktensor']
        if not getattr(mod, '__file__', '').startswith(reference_root):
            raise RuntimeError('a different sktensor (%s) is already imported; run the '
                               'reference in its own process' % mod.__file__)
        return mod
    import sktensor
    assert sktensor.__file__.startswith(reference_root), sktensor.__file__
    return sktensor

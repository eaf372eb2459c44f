"""CP decomposition by alternating least squares -- drop-in for ``sktensor/cp.py``.

Same signature, defaults, return tuple and quirks as the reference (cp.py:46-171); the
loop body runs on the GPU: one fused MTTKRP per mode (``skt_mttkrp_dense`` /
``skt_mttkrp_coo``), the Hadamard-of-Grams pseudo-inverse (``skt_hadamard_pinv``), the
factor update + normalisation + Gram refresh (``skt_solve_stats`` / ``skt_normalise_gram``)
and the fit (``skt_cp_fit``).  The tensor is uploaded once per call; per iteration the host
reads back two scalars (fit, pinv status).  With ``torch.distributed`` initialised the
tensor is sharded along its largest mode (see ``_parallel``).
"""
import logging
import time

import numpy as np

from . import _parallel
from .core import nvecs
from .ktensor import ktensor

_log = logging.getLogger('CP')
_DEF_MAXITER = 500
_DEF_INIT = 'nvecs'
_DEF_CONV = 1e-5
_DEF_FIT_METHOD = 'full'
_DEF_TYPE = float

__all__ = ['als', 'opt', 'wopt']


def als(X, rank, **kwargs):
    """Alternating least squares for the rank-``rank`` CP model of ``X``.

    Parameters
    ----------
    X : dtensor or sptensor
    rank : int
    init : {'nvecs', 'random'} or list of factor matrices (``[0]`` may be None), default 'nvecs'
    max_iter : int, default 500
    fit_method : {'full', None}, default 'full' -- ``None`` skips the fit and runs ``max_iter`` sweeps
    conv : float, default 1e-5 -- stop when the fit changes by less than this
    dtype : kept for signature compatibility; all arithmetic is FP64

    Returns
    -------
    P : ktensor, fit : float, itr : int (index of the last iteration), exectimes : ndarray
        (wall-clock seconds per iteration, device-synchronised; the reference reported
        process CPU time from the long-removed ``time.clock``)
    """
    ainit = kwargs.pop('init', _DEF_INIT)
    maxiter = kwargs.pop('max_iter', _DEF_MAXITER)
    fit_method = kwargs.pop('fit_method', _DEF_FIT_METHOD)
    conv = kwargs.pop('conv', _DEF_CONV)
    dtype = kwargs.pop('dtype', _DEF_TYPE)
    if not len(kwargs) == 0:
        raise ValueError('Unknown keywords (%s)' % (kwargs.keys()))

    if int(maxiter) < 1:
        # the reference leaves its loop without ever assigning the model and fails on an unbound name (cp.py:138,163-171)
        raise ValueError('max_iter must be at least 1 (got %s)' % str(maxiter))
    N = X.ndim
    # one upload per call: the 'nvecs' initialisation and the sweep engine share the device copy
    pinned_here = hasattr(X, 'cache_on_device') and getattr(X, '_dev', None) is None and ainit == 'nvecs'
    if pinned_here:
        X.cache_on_device()
    try:
        U = _init(ainit, X, N, rank, dtype)
        small = _run_small(X, rank, U, maxiter, conv, fit_method)
        if small is not None:
            return small
        eng = _parallel.CpAlsEngine(X, rank, U)
    finally:
        if pinned_here:
            X.drop_device_cache()

    try:
        return _sweeps(eng, U, maxiter, fit_method, conv)
    finally:
        eng.release()                     # gives the reserved SM back even when a sweep raises


def _sweeps(eng, U, maxiter, fit_method, conv):
    fit = 0
    exectimes = []
    itr = -1
    for itr in range(maxiter):
        tic = time.perf_counter()
        fitold = fit
        eng.sweep(first_iter=(itr == 0), want_fit=(fit_method == 'full'))
        eng.begin_read()
        if itr + 1 < maxiter:
            eng.prefetch_next_sweep()     # keeps the GPU busy across the host's convergence check
        if fit_method == 'full':
            fit = eng.read_fit()          # synchronises; raises ValueError on a non-finite Gram
        else:
            eng.check_status()
            fit = itr
        fitchange = abs(fitold - fit)
        exectimes.append(time.perf_counter() - tic)
        _log.debug('[%3d] fit: %.5f | delta: %7.1e | secs: %.5f' % (itr, fit, fitchange, exectimes[-1]))
        if itr > 0 and fitchange < conv:
            break

    lmbda = eng.download(U)               # fills the caller-visible list in place (cp.py:154,195-196)
    P = ktensor(U, lmbda)
    return P, fit, itr, np.array(exectimes)


_RUN_LOOP_MAX_ELEMS = 1 << 21


def _run_small(X, rank, U, maxiter, conv, fit_method):
    """Small dense problems are launch-latency bound: the whole loop runs behind one C call (``skt_cp_als_run``: one
    MTTKRP and one fused update launch per mode, one 24-byte read-back per iteration).  Returns None when the problem
    does not qualify (large or sharded tensors, sparse tensors, tall factors) -- those take the sweep engine."""
    import os
    from .dtensor import dtensor
    if not isinstance(X, dtensor) or os.environ.get('SKT_NO_RUN_LOOP', '0') == '1' or int(maxiter) < 1:
        return None
    if _parallel._Comm().active or X.size > _RUN_LOOP_MAX_ELEMS or X.size == 0:
        return None
    from . import _device as dev
    from . import _kernels as K
    dev.require_cuda()
    shape = tuple(int(s) for s in X.shape)
    R = int(rank)
    if R < 1:
        raise ValueError('rank must be a positive integer')
    if R > 64 or not K.cp_als_run_fits(shape, R):
        return None
    Ud, mask = [], 0
    for m in range(X.ndim):
        if U[m] is None:
            Ud.append(dev.empty((shape[m], R)))
            continue
        Um = np.asarray(U[m], dtype=np.float64)
        if Um.ndim != 2 or Um.shape != (shape[m], R):
            raise ValueError('Dimension mismatch of factor matrices')
        Ud.append(dev.to_device(Um))
        mask |= 1 << m
    try:
        lam, fit, itr, ex = K.cp_als_run(X._device_tensor(), shape, R, Ud, mask, int(maxiter), float(conv), fit_method == 'full')
    except ValueError as e:
        if 'infs or NaNs' in str(e):
            raise ValueError('array must not contain infs or NaNs')
        raise
    for m in range(X.ndim):
        U[m] = dev.to_host(Ud[m])
    if fit_method != 'full':
        fit = itr
    return ktensor(U, dev.to_host(lam)), fit, itr, ex


def opt(X, rank, **kwargs):
    """Placeholder kept for API parity: the reference's ``opt`` only parses its options (cp.py:174-183)."""
    kwargs.pop('init', _DEF_INIT)
    kwargs.pop('maxIter', _DEF_MAXITER)
    kwargs.pop('conv', _DEF_CONV)
    kwargs.pop('dtype', _DEF_TYPE)
    if not len(kwargs) == 0:
        raise ValueError('Unknown keywords (%s)' % (kwargs.keys()))


def wopt(X, rank, **kwargs):
    raise NotImplementedError()


def _init(init, X, N, rank, dtype):
    """Initial factors (cp.py:190-205): ``U[0]`` stays None; 'random' draws from the global legacy RNG."""
    Uinit = [None for _ in range(N)]
    if isinstance(init, list):
        Uinit = init
    elif init == 'random':
        for n in range(1, N):
            Uinit[n] = np.array(np.random.rand(X.shape[n], rank), dtype=dtype)
    elif init == 'nvecs':
        for n in range(1, N):
            Uinit[n] = np.array(nvecs(X, n, rank), dtype=dtype)
    else:
        raise ValueError('Unknown option (init=%s)' % str(init))
    return Uinit

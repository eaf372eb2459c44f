"""Tucker decomposition by higher-order orthogonal iteration -- drop-in for ``sktensor/tucker.py``.

Same signature, defaults (note ``maxIter``, not ``max_iter``) and return value ``(core, U)``
as the reference (tucker.py:36-123).  For dense tensors the whole loop stays on the device:
the tensor is uploaded once, every TTM of the chain is one ``skt_ttm_dense`` call on the
free (L, I_n, T) view (no transpose/reshape copies), the Gram of the small projected
tensor comes from ``skt_unfold_gram`` and its leading eigenvectors from the cluster
eigensolver ``skt_symeig_top``; the factors stay on the device between modes and the host
reads one scalar (the core's norm) per iteration.
"""
import logging
import time

import numpy as np

from .core import device_nvecs, flipsign, norm, nvecs, ttm
from .dtensor import dtensor
from .pyutils import is_number

__all__ = ['hooi', 'hosvd']

_log = logging.getLogger('TUCKER')
_DEF_MAXITER = 500
_DEF_INIT = 'nvecs'
_DEF_CONV = 1e-7


def hooi(X, rank, **kwargs):
    """Tucker decomposition of ``X`` with n-rank ``rank`` (list, or one number for all modes).

    Parameters: ``init`` ('nvecs' default, 'random', or a list of factors), ``maxIter``
    (default 500), ``conv`` (default 1e-7), ``dtype``.  Returns ``(core, U)``.
    """
    ainit = kwargs.pop('init', _DEF_INIT)
    maxIter = kwargs.pop('maxIter', _DEF_MAXITER)
    conv = kwargs.pop('conv', _DEF_CONV)
    dtype = kwargs.pop('dtype', X.dtype)
    if not len(kwargs) == 0:
        raise ValueError('Unknown keywords (%s)' % (kwargs.keys()))

    ndims = X.ndim
    if is_number(rank):
        rank = [rank] * ndims
    rank = [int(r) for r in rank]          # the reference carries floats here (tucker.py:90-91)

    if isinstance(X, dtensor):
        # one upload per call: the 'nvecs' initialisation (hosvd) and the iteration share the device copy
        pinned_here = getattr(X, '_dev', None) is None
        if pinned_here:
            X.cache_on_device()
        try:
            U = _init(ainit, X, ndims, rank, dtype)
            core, fit = _hooi_dense_device(X, rank, U, maxIter, conv)
        finally:
            if pinned_here:
                X.drop_device_cache()
        return core, U
    U = _init(ainit, X, ndims, rank, dtype)

    # generic route for other tensor types: the reference's loop on their own ttm / nvecs
    normX = norm(X)
    fit = 0
    for itr in range(maxIter):
        tic = time.perf_counter()
        fitold = fit
        for n in range(ndims):
            Utilde = ttm(X, U, n, transp=True, without=True)
            U[n] = nvecs(Utilde, n, rank[n])
        core = ttm(Utilde, U, n, transp=True)
        fit = 1 - (np.sqrt(normX ** 2 - norm(core) ** 2) / normX)
        fitchange = abs(fitold - fit)
        _log.debug('[%3d] fit: %.5f | delta: %7.1e | secs: %.5f' % (itr, fit, fitchange, time.perf_counter() - tic))
        if itr > 1 and fitchange < conv:
            break
    return core, U


def _hooi_dense_device(X, rank, U, maxIter, conv):
    from . import _device as dev
    from . import _kernels as K
    t = dev.torch()
    N = X.ndim
    shape = tuple(int(s) for s in X.shape)
    Xd = X._device_tensor()
    normX2 = float(dev.to_host(K.sumsq(Xd))[0])
    normX = np.sqrt(normX2)
    Ud = [None if u is None else dev.to_device(u) for u in U]
    fit = 0
    core_d, core_shape = None, None
    for itr in range(maxIter):
        tic = time.perf_counter()
        fitold = fit
        for n in range(N):
            # Utilde = X x_{m != n} U_m^T (tucker.py:103, core.py:99-102).  Only a contraction over the LAST mode
            # (K-major GEMM, mode order kept) or over the FIRST mode (M-major GEMM, result stored with the new mode last)
            # runs on the TMA / FP64 tensor-core kernel, so the chain is ordered greedily to hit those two positions:
            # `layout` lists which original mode sits at each position; a mode stuck in the middle (one contraction of
            # the n = 0 chain of a 3-way tensor) takes the tiled GEMM; the Gram works on any position.
            Yd, yshape, layout = Xd, shape, list(range(N))
            todo = [m for m in range(N) if m != n]
            while todo:
                if layout[-1] in todo:
                    m = layout[-1]
                    Yd, yshape = K.ttm_dense(Yd, yshape, Ud[m], len(layout) - 1, True)
                elif layout[0] in todo:
                    m = layout[0]
                    Yd, yshape = K.ttm_dense_first_rot(Yd, yshape, Ud[m])
                    layout = layout[1:] + [m]
                else:
                    m = todo[0]
                    Yd, yshape = K.ttm_dense(Yd, yshape, Ud[m], layout.index(m), True)
                todo.remove(m)
            # U_n = leading eigenvectors of Utilde_(n) Utilde_(n)^T, descending, sign-fixed (core.py:275-306)
            G = K.unfold_gram(Yd, yshape, layout.index(n))
            Ud[n] = device_nvecs(G, rank[n])
        # core = Utilde x_{N-1} U_{N-1}^T (tucker.py:105); the last mode sits first after its own chain
        if layout[0] == N - 1:
            core_d, core_shape = K.ttm_dense_first_rot(Yd, yshape, Ud[N - 1])
            layout = layout[1:] + [N - 1]
        else:
            core_d, core_shape = K.ttm_dense(Yd, yshape, Ud[N - 1], layout.index(N - 1), True)
        normcore2 = float(dev.to_host(K.sumsq(core_d))[0])
        fit = 1 - (np.sqrt(max(normX2 - normcore2, 0.0)) / normX)
        fitchange = abs(fitold - fit)
        _log.debug('[%3d] fit: %.5f | delta: %7.1e | secs: %.5f' % (itr, fit, fitchange, time.perf_counter() - tic))
        if itr > 1 and fitchange < conv:
            break
    for n in range(N):                       # the caller's list is filled in place, like the reference's U[n] = ...
        if Ud[n] is not None:
            U[n] = np.array(dev.to_host(Ud[n]), dtype=float)
    # back to the reference's mode order (the last sweep's chain left the modes rotated)
    core = np.ascontiguousarray(np.transpose(dev.to_host(core_d), [layout.index(m) for m in range(N)]))
    return dtensor(core), fit


def hosvd(X, rank, dims=None, dtype=None, compute_core=True):
    """Truncated higher-order SVD: ``nvecs`` per mode, optionally the core (tucker.py:125-137)."""
    U = [None for _ in range(X.ndim)]
    if dims is None:
        dims = range(X.ndim)
    if dtype is None:
        dtype = X.dtype
    for d in dims:
        U[d] = np.array(nvecs(X, d, rank[d]), dtype=dtype)
    if compute_core:
        return U, X.ttm(U, transp=True)
    return U


def _init(init, X, N, rank, dtype):
    """Initial factors; the first one is left to the first sweep (tucker.py:139-150)."""
    if isinstance(init, list):
        return init
    if init == 'random':
        return [None] + [np.array(np.random.rand(X.shape[n], rank[n]), dtype=dtype) for n in range(1, N)]
    if init == 'nvecs':
        return hosvd(X, rank, range(1, N), dtype=dtype, compute_core=False)
    raise ValueError('Unknown option (init=%s)' % str(init))

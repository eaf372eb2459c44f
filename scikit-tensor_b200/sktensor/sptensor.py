"""Sparse COO tensors -- host-side mirror of ``sktensor/sptensor.py`` backed by sm_100a kernels.

The container semantics of the reference are kept (tuple of subscript arrays + values,
duplicates allowed, ``accumfun`` de-duplication, inferred shape).  ``uttkrp``, ``ttv`` with
at most one surviving mode, and ``norm`` run on the device: the non-zeros are uploaded once,
radix-sorted per output mode by ``skt_coo_build`` and cached on the object.  The remaining
re-shaping helpers (``unfold``/``fold``/``transpose``/``concatenate``/indexing) are plain
host bookkeeping and not part of the accelerated path.
"""
import numpy as np
from scipy.sparse import coo_matrix

from .core import tensor_mixin
from .dtensor import unfolded_dtensor
from .pyutils import inherit_docstring_from
from .utils import accum

__all__ = ['concatenate', 'fromarray', 'sptensor', 'unfolded_sptensor']


class sptensor(tensor_mixin):
    """Sparse tensor in coordinate format (reference: sptensor.py:36-287).

    Parameters
    ----------
    subs : tuple of index arrays, one per mode
    vals : values of the stored entries
    shape : optional; inferred as ``max + 1`` per mode when absent
    dtype : optional value dtype
    accumfun : optional reducer applied to the values of duplicate coordinates

    >>> S = sptensor(([0, 1, 2], [3, 2, 0], [2, 2, 2]), [1, 1, 1], shape=(10, 20, 5), dtype=float)
    >>> S.shape
    (10, 20, 5)
    """

    def __init__(self, subs, vals, shape=None, dtype=None, accumfun=None, issorted=False):
        if not isinstance(subs, tuple):
            raise ValueError('Subscripts must be a tuple of array-likes')
        if len(subs[0]) != len(vals):
            raise ValueError('Subscripts and values must be of equal length')
        if dtype is None:
            dtype = np.array(vals).dtype
        for s in subs:
            if np.array(s).dtype.kind != 'i':
                raise ValueError('Subscripts must be integers')
        vals = np.array(vals, dtype=dtype)
        if accumfun is not None:
            vals, subs = _accumulate(subs, vals, accumfun)
        self.subs = subs
        self.vals = vals
        self.dtype = dtype
        self.issorted = issorted
        self.accumfun = accumfun
        if shape is None:
            self.shape = tuple(int(np.max(s)) + 1 for s in subs)
        else:
            self.shape = tuple(int(d) for d in shape)
        self.ndim = len(subs)
        self._coo = None

    # ------------------------------------------------------------------ device residency
    def _device_coo(self):
        """Sorted device copies (one per output mode), built on first use and cached."""
        if self._coo is None:
            from . import _kernels as K
            self._coo = K.CooHandle([np.asarray(s) for s in self.subs], self.vals, self.shape)
        return self._coo

    def drop_device_cache(self):
        if self._coo is not None:
            self._coo.close()
        self._coo = None
        return self

    # ------------------------------------------------------------------ the hot path
    @inherit_docstring_from(tensor_mixin)
    def uttkrp(self, U, mode):
        # reference: sptensor.py:216-229 -- R ttv passes; here all R columns in one kernel
        from . import _device as dev
        from . import _kernels as K
        mode = int(mode)
        if not (0 <= mode < self.ndim):
            raise ValueError('Invalid dimensions')
        R = U[1].shape[1] if mode == 0 else U[0].shape[1]
        Ud = []
        for m in range(self.ndim):
            if m == mode:
                Ud.append(None)
                continue
            Um = np.asarray(U[m])
            if Um.ndim != 2 or Um.shape[0] != self.shape[m]:
                raise ValueError('Multiplicant is wrong size')
            if Um.shape[1] != R:
                raise ValueError('All matrices must have same number of columns')
            Ud.append(dev.to_device(Um))
        M = K.mttkrp_coo(self._device_coo(), Ud, R, mode)
        return dev.to_host(M)

    def _ttv_compute(self, v, dims, vidx, remdims):
        # reference: sptensor.py:153-177
        from . import _device as dev
        from . import _kernels as K
        remdims = [int(r) for r in remdims]
        if len(remdims) > 1:
            # more than one surviving mode: scale the values on the host and merge duplicates
            w = np.asarray(self.vals, dtype=float)
            for d, k in zip(dims, vidx):
                w = w * np.asarray(v[k])[self.subs[d]]
            return sptensor(tuple(self.subs[r] for r in remdims), w,
                            shape=tuple(self.shape[r] for r in remdims), accumfun=np.sum)
        cols = [None] * self.ndim
        for d, k in zip(dims, vidx):
            cols[int(d)] = dev.to_device(np.asarray(v[k], dtype=np.float64).reshape(-1, 1))
        coo = self._device_coo()
        if not remdims:
            one = dev.to_device(np.ones(1))
            return float(dev.to_host(K.coo_innerprod(coo, cols, one, 1))[0])
        r = remdims[0]
        c = dev.to_host(K.mttkrp_coo(coo, cols, 1, r)).ravel()
        (nz,) = c.nonzero()
        return sptensor((nz,), c[nz], (self.shape[r],))

    def norm(self):
        """Frobenius norm of the stored values; duplicates are not merged (sptensor.py:274-282)."""
        from . import _device as dev
        from . import _kernels as K
        if len(self.vals) == 0:
            return 0.0
        return float(np.sqrt(dev.to_host(K.coo_sumsq(self._device_coo()))[0]))

    # ------------------------------------------------------------------ host bookkeeping (off the path)
    def __eq__(self, other):
        if isinstance(other, sptensor):
            self._sort()
            other._sort()
            return (len(self.vals) == len(other.vals) and (self.vals == other.vals).all()
                    and (np.array(self.subs) == np.array(other.subs)).all())
        if isinstance(other, np.ndarray):
            return (self.toarray() == other).all()
        raise NotImplementedError('Unsupported object class for sptensor.__eq__ (%s)' % type(other))

    __hash__ = None

    def __getitem__(self, idx):
        if len(idx) != self.ndim:
            raise ValueError('subscripts must be complete')
        hit = np.ones(len(self.vals), dtype=bool)
        for m in range(self.ndim):
            hit &= (np.asarray(self.subs[m]) == idx[m])
        found = self.vals[hit]
        if len(found) == 0:
            return 0
        if len(found) > 1:
            if self.accumfun is None:
                raise ValueError('Duplicate entries without specified accumulation function')
            return self.accumfun(found)
        return found

    def __sub__(self, other):
        if not isinstance(other, np.ndarray):
            raise NotImplementedError()
        res = -other
        res[self.subs] += self.vals
        return res

    def _sort(self):
        order = np.lexsort(np.array(self.subs))
        self.subs = tuple(np.asarray(s)[order] for s in self.subs)
        self.vals = self.vals[order]
        self.issorted = True
        self.drop_device_cache()

    def _ttm_compute(self, V, mode, transp):
        """Sparse tensor times matrix (sptensor.py:139-151): the result is dense in the multiplied mode.

        ``Y[.., j, ..] = sum_i X[.., i, ..] V[j, i]`` is a segmented reduction over the non-zeros grouped by their
        other subscripts -- the sparse MTTKRP kernel on the 2-way view (row = linear index of the other modes,
        column = i_mode) with the single "factor" ``V^T``: no unfolding matrix, no SciPy product.  The dense result
        is returned as a ``dtensor`` whose device copy is kept, so the rest of a TTM chain (tucker.hooi on a sparse
        tensor, tucker.py:103) continues on the device without a re-upload."""
        from . import _device as dev
        from . import _kernels as K
        from .dtensor import dtensor
        V = np.asarray(V, dtype=np.float64)
        mode = int(mode)
        if V.ndim != 2 or (V.shape[0] if transp else V.shape[1]) != self.shape[mode]:
            raise ValueError('shapes %s and %s not aligned' % (V.shape, self.shape))
        Vt = V if transp else V.T                       # (I_mode x p): row i holds V[:, i]
        p = int(Vt.shape[1])
        others = [m for m in range(self.ndim) if m != mode]
        oshape = tuple(self.shape[m] for m in others)
        nrows = int(np.prod(oshape)) if others else 1
        if nrows >= 2 ** 31 - 1:
            raise ValueError('the dense result of this sparse ttm (%d x %d) is too large' % (nrows, p))
        if len(self.vals) == 0:
            shape = list(self.shape)
            shape[mode] = p
            return dtensor(np.zeros(shape))
        rows = np.ravel_multi_index(tuple(np.asarray(self.subs[m]) for m in others), oshape) if others else \
            np.zeros(len(self.vals), dtype=np.int64)
        coo2 = K.CooHandle([np.asarray(rows, dtype=np.int64), np.asarray(self.subs[mode], dtype=np.int64)],
                           np.asarray(self.vals, dtype=np.float64), (nrows, self.shape[mode]), modes_mask=1)
        if p <= 128:
            M = K.mttkrp_coo(coo2, [None, dev.to_device(np.ascontiguousarray(Vt))], p, 0)
        else:                                           # the sparse kernel holds at most 128 columns per pass
            t = dev.torch()
            M = t.cat([K.mttkrp_coo(coo2, [None, dev.to_device(np.ascontiguousarray(Vt[:, c0:c0 + 128]))],
                                    min(128, p - c0), 0) for c0 in range(0, p, 128)], dim=1)
        coo2.close()
        Yd = M.reshape(oshape + (p,))
        if mode != self.ndim - 1:                       # the new mode goes back to its position (a device copy)
            perm = list(range(mode)) + [self.ndim - 1] + list(range(mode, self.ndim - 1))
            Yd = Yd.permute(perm).contiguous()
        Y = dtensor(dev.to_host(Yd))
        Y._dev = Yd
        return Y

    def unfold(self, rdims, cdims=None, transp=False):
        """Matricise: modes ``rdims`` index the rows, ``cdims`` the columns (sptensor.py:197-214)."""
        if isinstance(rdims, (int, np.integer)):
            rdims = [int(rdims)]
        if transp:
            cdims = rdims
            rdims = np.setdiff1d(range(self.ndim), cdims)[::-1]
        elif cdims is None:
            cdims = np.setdiff1d(range(self.ndim), rdims)[::-1]
        if sorted(list(rdims) + list(cdims)) != list(range(self.ndim)):
            raise ValueError('Incorrect specification of dimensions (rdims: %s, cdims: %s)'
                             % (str(rdims), str(cdims)))
        M = int(np.prod([self.shape[r] for r in rdims]))
        N = int(np.prod([self.shape[c] for c in cdims]))
        ridx = _build_idx(self.subs, self.vals, rdims, self.shape)
        cidx = _build_idx(self.subs, self.vals, cdims, self.shape)
        return unfolded_sptensor((self.vals, (ridx, cidx)), (M, N), rdims, cdims, self.shape)

    @inherit_docstring_from(tensor_mixin)
    def transpose(self, axes=None):
        if axes is None:
            raise NotImplementedError('Sparse tensor transposition without axes argument is not supported')
        return sptensor(tuple(self.subs[a] for a in axes), self.vals, [self.shape[a] for a in axes])

    def concatenate(self, tpl, axis=None):
        """Concatenate ``tpl[1:]`` onto this tensor along ``axis`` (sptensor.py:254-272)."""
        if axis is None:
            raise NotImplementedError('Sparse tensor concatenation without axis argument is not supported')
        T = self
        for other in tpl[1:]:
            T = _single_concatenate(T, other, axis=axis)
        return T

    def toarray(self):
        """Dense ndarray; for duplicate coordinates the last one wins (sptensor.py:284-287)."""
        A = np.zeros(self.shape)
        A.put(np.ravel_multi_index(self.subs, tuple(self.shape)), self.vals)
        return A


class unfolded_sptensor(coo_matrix):
    """Sparse matrix that remembers the tensor it came from (sptensor.py:290-359)."""

    def __init__(self, tpl, shape, rdims, cdims, ten_shape, dtype=None, copy=False):
        self.ten_shape = np.array(ten_shape)
        if isinstance(rdims, (int, np.integer)):
            rdims = [int(rdims)]
        if cdims is None:
            cdims = np.setdiff1d(range(len(self.ten_shape)), rdims)[::-1]
        self.rdims = [int(r) for r in rdims]
        self.cdims = [int(c) for c in cdims]
        if shape is None:
            shape = (int(np.prod(self.ten_shape[self.rdims])) if len(self.rdims) else 1,
                     int(np.prod(self.ten_shape[self.cdims])) if len(self.cdims) else 1)
        super(unfolded_sptensor, self).__init__(tpl, shape=shape, dtype=dtype, copy=copy)

    def fold(self):
        """Back to an ``sptensor`` of shape ``ten_shape``."""
        N = len(self.ten_shape)
        subs = [None] * N
        if self.rdims:
            for m, ix in zip(self.rdims, np.unravel_index(self.row, tuple(self.ten_shape[self.rdims]))):
                subs[m] = np.asarray(ix, dtype=np.int64)
        if self.cdims:
            for m, ix in zip(self.cdims, np.unravel_index(self.col, tuple(self.ten_shape[self.cdims]))):
                subs[m] = np.asarray(ix, dtype=np.int64)
        return sptensor(tuple(subs), self.data, tuple(int(s) for s in self.ten_shape))


_DEVICE_ACCUM_MIN = 4096       # below this, sorting a handful of coordinates is host bookkeeping, not a kernel's job


def _accumulate(subs, vals, func):
    """Merge duplicate coordinates (sptensor.py:80-84 -> utils.accum).  The sum / max / min reducers run on the
    device (radix sort of the linear keys + one pass over the runs, ``skt_coo_accumulate_host``); arbitrary Python
    callables cannot, and keep the reference's host loop."""
    op = None
    if func in (np.sum, sum):
        op = 'sum'
    elif func in (np.max, np.amax, max):
        op = 'max'
    elif func in (np.min, np.amin, min):
        op = 'min'
    subs = tuple(np.asarray(s) for s in subs)
    if op is None or len(vals) < _DEVICE_ACCUM_MIN or any(s.dtype.kind != 'i' for s in subs) or \
            np.asarray(vals).dtype.kind not in 'fiub':
        return accum(subs, vals, issorted=False, with_subs=True, func=func)
    from . import _kernels as K
    return K.coo_accumulate(subs, vals, op)


def fromarray(A):
    """Sparse tensor holding the non-zero entries of a dense array (sptensor.py:362-366)."""
    A = np.asarray(A)
    subs = np.nonzero(A)
    return sptensor(subs, A[subs], shape=A.shape, dtype=A.dtype)


def concatenate(tpl, axis=None):
    return tpl[0].concatenate(tpl, axis=axis)


def _single_concatenate(ten, other, axis):
    if len(ten.shape) != len(other.shape):
        raise ValueError("len(tshape) != len(oshape")
    for m in range(len(ten.shape)):
        if m != axis and ten.shape[m] != other.shape[m]:
            raise ValueError("Dimensions must match")
    subs = []
    for m in range(len(ten.shape)):
        shift = ten.shape[axis] if m == axis else 0
        subs.append(np.concatenate((np.asarray(ten.subs[m]), np.asarray(other.subs[m]) + shift)))
    shape = list(ten.shape)
    shape[axis] = ten.shape[axis] + other.shape[axis]
    return sptensor(tuple(subs), np.concatenate((ten.vals, other.vals)), shape)


def _build_idx(subs, vals, dims, tshape):
    dims = np.array(dims, ndmin=1)
    if len(dims) == 0:
        return np.zeros(len(vals), dtype=np.int64)
    return np.ravel_multi_index(tuple(subs[int(d)] for d in dims), tuple(tshape[int(d)] for d in dims))

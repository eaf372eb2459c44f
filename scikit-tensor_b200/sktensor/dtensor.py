"""Dense tensors -- host-side mirror of ``sktensor/dtensor.py`` backed by sm_100a kernels.

``dtensor`` is still a zero-copy ``numpy.ndarray`` view (any dtype, any strides); the
numerical methods upload it once (C-order FP64, Appendix A.15/16 of SURVEY.md) and call the
C ABI.  ``cp_als`` / ``tucker_hooi`` keep that device copy for the whole call; standalone
method calls re-upload unless the caller opts into ``cache_on_device()``.
"""
import numpy as np

from .core import tensor_mixin
from .pyutils import inherit_docstring_from

__all__ = ['dtensor', 'unfolded_dtensor']


class dtensor(tensor_mixin, np.ndarray):
    """Dense tensor (reference: dtensor.py:29-171).

    >>> T = np.zeros((3, 4, 2))
    >>> T = dtensor(T)
    """

    def __new__(cls, input_array):
        obj = np.asarray(input_array).view(cls)
        obj._dev = None
        return obj

    def __array_finalize__(self, obj):
        self._dev = None            # views and results never inherit a device copy

    def __array_wrap__(self, out_arr, context=None, return_scalar=False):
        if out_arr.ndim == 0:
            return out_arr[()]
        return np.asarray(out_arr).view(type(self))

    def __eq__(self, other):
        return np.equal(self, other)

    def __ne__(self, other):
        return np.not_equal(self, other)

    __hash__ = None

    # ------------------------------------------------------------------ device residency
    def _device_tensor(self):
        """C-order FP64 copy in HBM (cached only after ``cache_on_device``)."""
        if self._dev is not None:
            return self._dev
        from . import _device as dev
        return dev.to_device(np.asarray(self))

    def cache_on_device(self):
        """Upload now and keep the copy for later method calls.

        The cache is NOT invalidated by in-place writes to the array (NumPy offers no
        hook); call ``drop_device_cache()`` after modifying the data.
        """
        from . import _device as dev
        self._dev = dev.to_device(np.asarray(self))
        return self

    def drop_device_cache(self):
        self._dev = None
        return self

    # ------------------------------------------------------------------ the hot path
    @inherit_docstring_from(tensor_mixin)
    def uttkrp(self, U, n):
        # reference: dtensor.py:163-167 (unfold + khatrirao + dot); here one fused kernel
        from . import _device as dev
        from . import _kernels as K
        n = int(n)
        if not (0 <= n < self.ndim):
            raise ValueError('Invalid dimensions')
        rank = None
        Ud = []
        for m in range(self.ndim):
            if m == n:
                Ud.append(None)
                continue
            Um = np.asarray(U[m])
            if Um.ndim != 2:
                raise ValueError('A must be a tuple of matrices (A[%d].ndim = %d)' % (m, Um.ndim))
            if rank is None:
                rank = Um.shape[1]
            elif Um.shape[1] != rank:
                raise ValueError('All matrices must have same number of columns')
            if Um.shape[0] != self.shape[m]:
                raise ValueError('shapes %s and %s not aligned' % (self.shape, Um.shape))
            Ud.append(dev.to_device(Um))
        M = K.mttkrp_dense(self._device_tensor(), self.shape, Ud, rank, n)
        return unfolded_dtensor(dev.to_host(M), n, self.shape)

    def norm(self):
        """Frobenius norm (dtensor.py:152-161), reduced on the device in a fixed order."""
        from . import _device as dev
        from . import _kernels as K
        return float(np.sqrt(dev.to_host(K.sumsq(self._device_tensor()))[0]))

    def _ttm_compute(self, V, mode, transp):
        # reference: dtensor.py:58-76 (transpose, reshape, dot, reshape, transpose)
        from . import _device as dev
        from . import _kernels as K
        V = np.asarray(V)
        mode = int(mode)
        need = V.shape[0] if transp else V.shape[1]
        if V.ndim != 2 or need != self.shape[mode]:
            raise ValueError('shapes %s and %s not aligned' % (V.shape, self.shape))
        Y, _ = K.ttm_dense(self._device_tensor(), self.shape, dev.to_device(V), mode, transp)
        return dtensor(dev.to_host(Y))

    def _ttv_compute(self, v, dims, vidx, remdims):
        # reference: dtensor.py:78-98.  A vector is a one-column matrix: chain TTMs (last mode
        # first keeps the tensor small), then drop the singleton modes.
        from . import _device as dev
        from . import _kernels as K
        if not isinstance(v, tuple):
            raise ValueError('v must be a tuple of vectors')
        Xd = self._device_tensor()
        shape = tuple(self.shape)
        for i in range(len(dims) - 1, -1, -1):
            col = dev.to_device(np.asarray(v[vidx[i]], dtype=np.float64).reshape(-1, 1))
            Xd, shape = K.ttm_dense(Xd, shape, col, int(dims[i]), True)
        out = dev.to_host(Xd).reshape([self.shape[d] for d in remdims])
        if out.ndim == 0:
            return out.reshape(1)      # the reference returns a 1-element array here
        return dtensor(out)

    def ttt(self, other, modes=None):
        pass

    # ------------------------------------------------------------------ layout helpers (host, no arithmetic)
    def unfold(self, mode):
        """Mode-``mode`` matricisation in Kolda-Bader column order (dtensor.py:103-150).

        Provided for callers that want the matrix; the accelerated path never builds it.

        >>> T = dtensor(np.arange(24.).reshape(3, 4, 2))
        >>> T.unfold(1).shape
        (4, 6)
        """
        N = self.ndim
        rest = [m for m in range(N - 1, -1, -1) if m != mode]
        arr = np.transpose(np.asarray(self), [mode] + rest)
        ncols = int(np.prod([self.shape[m] for m in rest])) if rest else 1
        return unfolded_dtensor(arr.reshape(self.shape[mode], ncols), mode, self.shape)

    @inherit_docstring_from(tensor_mixin)
    def transpose(self, axes=None):
        return dtensor(np.transpose(np.array(self), axes=axes))


class unfolded_dtensor(np.ndarray):
    """Matrix that remembers the tensor it was unfolded from (dtensor.py:174-194)."""

    def __new__(cls, input_array, mode, ten_shape):
        obj = np.asarray(input_array).view(cls)
        obj.ten_shape = ten_shape
        obj.mode = mode
        return obj

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self.ten_shape = getattr(obj, 'ten_shape', None)
        self.mode = getattr(obj, 'mode', None)

    def fold(self):
        shape = tuple(self.ten_shape)
        N = len(shape)
        rest = [m for m in range(N - 1, -1, -1) if m != self.mode]
        order = [self.mode] + rest
        arr = np.asarray(self).reshape([shape[m] for m in order])
        return dtensor(np.transpose(arr, np.argsort(order)))

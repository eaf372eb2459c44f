"""Tensor base class and multilinear helpers -- host-side mirror of ``sktensor/core.py``.

Same names, argument meaning and error messages as the reference; the arithmetic on the
CP-ALS / HOOI path (``uttkrp``, ``ttm``, ``ttv``, ``norm``, ``nvecs``' Gram) runs in
``libsktensor_b200.so``.  Pure index bookkeeping stays in Python.
"""
import sys

import numpy as np

from .pyutils import is_sequence

__all__ = [
    'tensor_mixin', 'istensor', 'norm', 'transpose', 'ttm', 'ttv', 'unfold',
    'check_multiplication_dims', 'innerprod', 'nvecs', 'flipsign', 'center_matrix',
    'khatrirao', 'teneye', 'tvecmat',
]


class tensor_mixin(object):
    """Base of ``dtensor`` and ``sptensor`` (reference: core.py:37-204)."""

    def ttm(self, V, mode=None, transp=False, without=False):
        """Tensor-times-matrix, for one matrix or a list of matrices (core.py:50-103).

        With a list, ``mode`` selects the modes to multiply (all when ``None``);
        ``without=True`` multiplies every mode *except* those.  ``transp=True`` applies
        the transposed matrices.  Modes are processed in ascending order.
        """
        if mode is None:
            mode = range(self.ndim)
        if isinstance(V, np.ndarray):
            return self._ttm_compute(V, mode, transp)
        if not is_sequence(V):
            raise ValueError('V must be a matrix or a sequence of matrices')
        dims, vidx = check_multiplication_dims(mode, self.ndim, len(V), vidx=True, without=without)
        Y = self
        for d, k in zip(dims, vidx):
            Y = Y._ttm_compute(V[k], d, transp)
        return Y

    def ttv(self, v, modes=[], without=False):
        """Tensor-times-vector(s) (core.py:105-127)."""
        if not isinstance(v, tuple):
            v = (v, )
        dims, vidx = check_multiplication_dims(modes, self.ndim, len(v), vidx=True, without=without)
        for d, k in zip(dims, vidx):
            if len(v[k]) != self.shape[d]:
                raise ValueError('Multiplicant is wrong size')
        remdims = np.setdiff1d(range(self.ndim), dims)
        return self._ttv_compute(v, dims, vidx, remdims)

    # implemented by the tensor classes
    def _ttm_compute(self, V, mode, transp):
        raise NotImplementedError

    def _ttv_compute(self, v, dims, vidx, remdims):
        raise NotImplementedError

    def unfold(self, rdims, cdims=None, transp=False):
        raise NotImplementedError

    def uttkrp(self, U, mode):
        """Matricised tensor times Khatri-Rao product (core.py:145-183):

        ``M[i_n, r] = sum X[i_0..i_{N-1}] * prod_{m != mode} U[m][i_m, r]``.
        ``U[mode]`` is ignored and may be ``None``.
        """
        raise NotImplementedError

    def transpose(self, axes=None):
        raise NotImplementedError


def istensor(X):
    return isinstance(X, tensor_mixin)


def _method_caller(name):
    def call(obj, *args, **kwargs):
        if not istensor(obj):
            raise ValueError('%s() object must be tensor (%s)' % (name, type(obj)))
        return getattr(obj, name)(*args, **kwargs)
    call.__name__ = name
    call.__doc__ = 'Module-level form of ``tensor.%s`` (core.py:211-240).' % name
    return call


for _name in ('norm', 'transpose', 'ttm', 'ttv', 'unfold'):
    setattr(sys.modules[__name__], _name, _method_caller(_name))


def check_multiplication_dims(dims, N, M, vidx=False, without=False):
    """Normalise "which operand multiplies which mode" (core.py:243-265).

    Returns the sorted mode list and, with ``vidx=True``, for each of them the position
    of its operand in the caller's sequence of ``M`` operands.
    """
    dims = np.array(dims, ndmin=1)
    if dims.size == 0:
        dims = np.arange(N)
    if without:
        dims = np.setdiff1d(np.arange(N), dims)
    if not np.isin(dims, np.arange(N)).all():
        raise ValueError('Invalid dimensions')
    order = np.argsort(dims)
    sdims = dims[order]
    if not vidx:
        return sdims
    P = len(dims)
    if M > N:
        raise ValueError('More multiplicants than dimensions')
    if M != N and M != P:
        raise ValueError('Invalid number of multiplicants')
    # one operand per selected mode -> positional; one operand per tensor mode -> indexed by mode
    return sdims, (order if P == M else sdims)


def innerprod(X, Y):
    """Flat inner product of two arrays (core.py:268-272)."""
    return np.dot(np.asarray(X).ravel(), np.asarray(Y).ravel())


def flipsign(U):
    """Flip column signs so each column's largest-magnitude entry is positive (core.py:297-306)."""
    rows = np.abs(U).argmax(axis=0)
    neg = U[rows, np.arange(U.shape[1])] < 0
    U[:, neg] = -U[:, neg]
    return U


def nvecs(X, n, rank, do_flipsign=True, dtype=float):
    """Leading ``rank`` eigenvectors of ``X_(n) X_(n)^T`` in descending order (core.py:275-294).

    Dense tensors: the I_n x I_n Gram is accumulated on the device straight from the
    C-order tensor (``skt_unfold_gram``; no unfolding copy) and its leading eigenvectors
    come from the cluster eigensolver ``skt_symeig_top`` (tridiagonalisation in distributed
    shared memory, Sturm multisection, inverse iteration; order and sign fixed on the
    device).  Modes longer than 2048 exceed that kernel's shared-memory design and take the
    library eigensolver instead.  Sparse tensors (the reference's ``eigsh`` route, core.py:280-283)
    apply ``X_(n) X_(n)^T`` through the sparse MTTKRP kernel on the 2-way view of the non-zeros:
    short modes assemble the Gram for the same eigensolver, long ones run a restarted block
    Lanczos iteration (``sktensor/_eigs.py``).
    """
    from .dtensor import dtensor
    rank = int(rank)
    if isinstance(X, dtensor):
        from . import _device as dev
        from . import _kernels as K
        Xd = X._device_tensor()
        G = K.unfold_gram(Xd, X.shape, n)
        return np.array(dev.to_host(device_nvecs(G, rank, do_flipsign)), dtype=float)
    from .sptensor import sptensor
    if isinstance(X, sptensor):
        from ._eigs import sparse_nvecs
        return sparse_nvecs(X, n, rank, do_flipsign)
    raise ValueError('nvecs() needs a dtensor or sptensor (%s)' % type(X))


def device_nvecs(G, rank, do_flipsign=True):
    """Leading eigenvectors of the symmetric device matrix ``G`` -> device (n x rank), descending, sign-fixed."""
    from . import _device as dev
    from . import _kernels as K
    n = int(G.shape[0])
    if n <= K.SYMEIG_MAX_N:
        U, info = K.symeig_top(G, rank, flipsign=do_flipsign)
        return U
    t = dev.torch()
    _, V = t.linalg.eigh(G)
    V = V[:, -rank:].flip(1).contiguous()
    if do_flipsign:
        rows = V.abs().argmax(dim=0)
        sgn = t.where(V[rows, t.arange(V.shape[1], device=V.device)] < 0, -1.0, 1.0)
        V = V * sgn
    return V


def center_matrix(X):
    """Subtract the column means (core.py:319-321)."""
    return X - X.mean(axis=0)


def khatrirao(A, reverse=False):
    """Column-wise Khatri-Rao product of a tuple of matrices (core.py:333-378).

    Row ``(i_0, ..., i_{k-1})`` (C-order, last matrix fastest; with ``reverse=True`` the
    first matrix fastest) holds ``prod_j A[j][i_j, :]``.  The accelerated MTTKRP never
    forms this matrix; the function exists for callers and tests that want it explicitly.
    """
    if not isinstance(A, tuple):
        raise ValueError('A must be a tuple of array likes')
    ncol = A[0].shape[1]
    for i, Ai in enumerate(A):
        if Ai.ndim != 2:
            raise ValueError('A must be a tuple of matrices (A[%d].ndim = %d)' % (i, Ai.ndim))
        if Ai.shape[1] != ncol:
            raise ValueError('All matrices must have same number of columns')
    mats = A[::-1] if reverse else A
    P = np.array(mats[0], dtype=A[0].dtype)
    for Ai in mats[1:]:
        P = (P[:, None, :] * Ai[None, :, :]).reshape(-1, ncol)
    return P.astype(A[0].dtype, copy=False)


def teneye(dim, order):
    """``order``-way identity tensor of side ``dim`` (intent of core.py:381-391)."""
    T = np.zeros((dim,) * order)
    idx = np.arange(dim)
    T[(idx,) * order] = 1
    return T


def tvecmat(m, n):
    """Commutation matrix ``K`` with ``K vec(A) = vec(A^T)`` for an m x n ``A`` (core.py:394-399)."""
    d = m * n
    perm = np.arange(d).reshape(m, n).T.ravel()
    K = np.zeros((d, d))
    K[np.arange(d), perm] = 1
    return K

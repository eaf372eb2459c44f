"""Kruskal tensors -- host-side mirror of ``sktensor/ktensor.py``.

``norm`` and ``innerprod`` (the two terms of the CP fit, cp.py:157-159) run on the device.
``innerprod`` uses one MTTKRP instead of the reference's rank-many full passes over ``X``:
``<P, X> = sum_r lambda_r sum_i U_{N-1}[i, r] * MTTKRP_{N-1}(X, U)[i, r]``.
"""
import numpy as np

from .dtensor import dtensor

__all__ = ['ktensor', 'vectorized_ktensor']


class ktensor(object):
    """Tensor held as a sum of ``rank`` outer products (reference: ktensor.py:27-184).

    Parameters
    ----------
    U : list of (I_m x rank) factor matrices -- kept by reference, not copied
    lmbda : optional weights, one per component (default: ones)
    """

    def __init__(self, U, lmbda=None):
        self.U = U
        self.shape = tuple(Ui.shape[0] for Ui in U)
        self.ndim = len(self.shape)
        self.rank = U[0].shape[1]
        self.lmbda = lmbda
        if any(Ui.shape[1] != self.rank for Ui in U):
            raise ValueError('Dimension mismatch of factor matrices')
        if lmbda is None:
            self.lmbda = np.ones(self.rank)

    def __eq__(self, other):
        if not isinstance(other, ktensor):
            raise NotImplementedError()
        if self.ndim != other.ndim or self.shape != other.shape:
            return False
        return (all((a == b).all() for a, b in zip(self.U, other.U))
                and (np.asarray(self.lmbda) == np.asarray(other.lmbda)).all())

    __hash__ = None

    def uttkrp(self, U, mode):
        """MTTKRP of a Kruskal tensor: ``U_mode (lambda * hadamard_{i != mode} U_i^T V_i)`` (ktensor.py:84-111)."""
        W = np.asarray(self.lmbda, dtype=float)[:, None] * np.ones((self.rank, U[1 if mode == 0 else 0].shape[1]))
        for i in range(self.ndim):
            if i != mode:
                W = W * np.dot(self.U[i].T, U[i])
        return np.dot(self.U[mode], W)

    def norm(self):
        """Frobenius norm from the Hadamard product of the factor Grams (ktensor.py:113-126)."""
        from . import _device as dev
        from . import _kernels as K
        R = self.rank
        G = dev.empty((self.ndim, R, R))
        for m in range(self.ndim):
            K.gram(dev.to_device(self.U[m]), out=G[m])
        lam = dev.to_device(np.asarray(self.lmbda, dtype=float))
        zero = dev.zeros((1,))
        one = dev.to_device(np.ones(1))
        out = K.cp_fit(G, lam, zero, one)       # out[1] = ||P||^2
        return float(np.sqrt(dev.to_host(out)[1]))

    def innerprod(self, X):
        """Inner product with a dense or sparse tensor (ktensor.py:128-150)."""
        from . import _device as dev
        from . import _kernels as K
        from .sptensor import sptensor
        R = self.rank
        lam = dev.to_device(np.asarray(self.lmbda, dtype=float))
        Ud = [dev.to_device(Um) for Um in self.U]
        for Um, s in zip(self.U, X.shape):
            if Um.shape[0] != s:
                raise ValueError('Multiplicant is wrong size')
        if isinstance(X, sptensor):
            return float(dev.to_host(K.coo_innerprod(X._device_coo(), Ud, lam, R))[0])
        if not isinstance(X, dtensor):
            X = dtensor(np.asarray(X))
        last = X.ndim - 1
        M = K.mttkrp_dense(X._device_tensor(), X.shape, Ud, R, last)
        return float(dev.to_host(K.weighted_dot(Ud[last], M, lam))[0])

    def toarray(self):
        """Dense ndarray of the represented tensor (ktensor.py:152-163): reconstructed on the device without the
        Khatri-Rao matrix (``skt_ktensor_full``), then copied back."""
        from . import _device as dev
        from . import _kernels as K
        Ud = [dev.to_device(np.asarray(Um, dtype=float)) for Um in self.U]
        lam = dev.to_device(np.asarray(self.lmbda, dtype=float))
        return dev.to_host(K.ktensor_full(Ud, lam, self.shape))

    def totensor(self):
        """Dense ``dtensor`` of the represented tensor (ktensor.py:165-175)."""
        return dtensor(self.toarray())

    def tovec(self):
        """All factor entries, factor by factor, in one vector (ktensor.py:177-184)."""
        v = np.concatenate([np.asarray(M).ravel() for M in self.U])
        return vectorized_ktensor(v, self.shape, self.lmbda)


class vectorized_ktensor(object):
    """Flat parameter vector of a ``ktensor`` (ktensor.py:187-203)."""

    def __init__(self, v, shape, lmbda):
        self.v = v
        self.shape = shape
        self.lmbda = lmbda

    def toktensor(self):
        rank = len(self.v) // sum(self.shape)
        U, off = [], 0
        for s in self.shape:
            U.append(self.v[off:off + s * rank].reshape((s, rank)))
            off += s * rank
        return ktensor(U, self.lmbda)

"""Thin Python wrappers: device tensors in, one C-ABI call each, device tensors out.

Everything here enqueues work on the current CUDA stream and returns without
synchronising.  Argument checking beyond what the ABI does itself is left to callers.
"""
import ctypes as C

import numpy as np

from . import _device as dev
from . import _lib


def _factor_ptrs(U, skip=None):
    return _lib.ptr_array([None if (u is None or m == skip) else u.data_ptr() for m, u in enumerate(U)])


# ------------------------------------------------------------------ dense MTTKRP
def mttkrp_dense(Xd, shape, U, rank, mode, out=None, naive=False):
    """``dtensor.uttkrp`` on device tensors (dtensor.py:163-167). ``U[mode]`` may be None."""
    lib = _lib.load()
    shp = _lib.shape_array(shape)
    nd = len(shape)
    if out is None:
        out = dev.empty((shape[mode], rank))
    ptrs = _factor_ptrs(U, mode)
    if naive:
        _lib.check(lib.skt_mttkrp_dense_naive(dev.ptr(Xd), shp, nd, ptrs, rank, mode, dev.ptr(out), dev.stream_ptr()))
        return out
    need = lib.skt_mttkrp_dense_workspace(shp, nd, rank, mode)
    if need == 0:      # the plan was rejected; the reason is in skt_last_error()
        _lib.check(-2)
    ws, wsz = dev.workspace('mttkrp').get(need)
    _lib.check(lib.skt_mttkrp_dense(dev.ptr(Xd), shp, nd, ptrs, rank, mode, dev.ptr(out), ws, wsz, dev.stream_ptr()))
    return out


def mttkrp_dense_host(X, U, rank, mode):
    """Host-buffer entry point (uploads, computes, downloads inside the C call)."""
    lib = _lib.load()
    dev.require_cuda()
    X = np.ascontiguousarray(X, dtype=np.float64)
    Uh = [None if (u is None or m == mode) else np.ascontiguousarray(u, dtype=np.float64) for m, u in enumerate(U)]
    M = np.empty((X.shape[mode], rank))
    ptrs = _lib.ptr_array([None if u is None else u.ctypes.data for u in Uh])
    _lib.check(lib.skt_mttkrp_dense_host(C.c_void_p(X.ctypes.data), _lib.shape_array(X.shape), X.ndim, ptrs,
                                         rank, mode, C.c_void_p(M.ctypes.data)))
    return M


def kr_reduce(T, dims, W, rank, mode, out=None):
    """Second stage of the dimension tree: MTTKRP of group mode ``mode`` from the partial tensor ``T``
    (``dims[0] x ... x dims[p-1] x rank``) and the group's factors ``W`` (``W[mode]`` ignored)."""
    lib = _lib.load()
    d = _lib.shape_array(dims)
    p = len(dims)
    if out is None:
        out = dev.empty((dims[mode], rank))
    need = lib.skt_kr_reduce_workspace(d, p, rank, mode)
    if need == 0:
        _lib.check(-2)
    ws, wsz = dev.workspace('mttkrp').get(need)
    _lib.check(lib.skt_kr_reduce(dev.ptr(T), d, p, _factor_ptrs(W, mode), rank, mode, dev.ptr(out), ws, wsz,
                                 dev.stream_ptr()))
    return out


def kr_reduce_slabs(T, dims, W, rank, mode, slabs):
    """``kr_reduce`` that leaves its per-split partial slabs in ``slabs`` (a device buffer of at least the workspace
    size); returns the number of slabs.  ``cp_update_cluster_slabs`` adds them while staging M."""
    lib = _lib.load()
    d = _lib.shape_array(dims)
    n = C.c_int(0)
    _lib.check(lib.skt_kr_reduce_slabs(dev.ptr(T), d, len(dims), _factor_ptrs(W, mode), rank, mode, dev.ptr(slabs),
                                       slabs.numel() * slabs.element_size(), C.byref(n), dev.stream_ptr()))
    return n.value


def kr_reduce_workspace(dims, rank, mode):
    return int(_lib.load().skt_kr_reduce_workspace(_lib.shape_array(dims), len(dims), rank, mode))


def ktensor_full(U, lam, shape):
    """Dense C-order reconstruction ``sum_r lam_r outer(U_0[:, r], ..., U_{N-1}[:, r])`` (ktensor.py:152-163)."""
    lib = _lib.load()
    out = dev.empty(tuple(int(s) for s in shape))
    R = int(U[0].shape[1])
    _lib.check(lib.skt_ktensor_full(_factor_ptrs(U), dev.ptr(lam), _lib.shape_array(shape), len(shape), R, dev.ptr(out),
                                    dev.stream_ptr()))
    return out


# ------------------------------------------------------------------ sparse
class CooHandle(object):
    """Owner of a device-resident sorted COO tensor (``skt_coo*``)."""

    def __init__(self, subs, vals, shape, modes_mask=0):
        lib = _lib.load()
        dev.require_cuda()
        self.shape = tuple(int(s) for s in shape)
        self.ndim = len(self.shape)
        self._subs = [np.ascontiguousarray(s, dtype=np.int64) for s in subs]
        self._vals = np.ascontiguousarray(vals, dtype=np.float64)
        self.nnz = int(self._vals.shape[0])
        h = C.c_void_p()
        ptrs = _lib.ptr_array([s.ctypes.data if self.nnz else None for s in self._subs])
        _lib.check(lib.skt_coo_build_host(ptrs, C.c_void_p(self._vals.ctypes.data) if self.nnz else None,
                                          self.nnz, _lib.shape_array(self.shape), self.ndim, modes_mask,
                                          C.byref(h)))
        self.h = h
        self._subs = self._vals = None   # host copies are not retained

    @classmethod
    def from_device(cls, subs_d, vals_d, shape, modes_mask=0):
        """Build from int64 subscript / float64 value tensors that already live on the device (``skt_coo_build``)."""
        lib = _lib.load()
        dev.require_cuda()
        self = cls.__new__(cls)
        self.shape = tuple(int(s) for s in shape)
        self.ndim = len(self.shape)
        self.nnz = int(vals_d.shape[0])
        self._subs = self._vals = None
        h = C.c_void_p()
        ptrs = _lib.ptr_array([s.data_ptr() if self.nnz else None for s in subs_d])
        _lib.check(lib.skt_coo_build(ptrs, dev.ptr(vals_d) if self.nnz else None, self.nnz, _lib.shape_array(self.shape),
                                     self.ndim, modes_mask, dev.stream_ptr(), C.byref(h)))
        self.h = h
        return self

    def device_bytes(self):
        return int(_lib.load().skt_coo_device_bytes(self.h))

    def close(self):
        if getattr(self, 'h', None):
            _lib.load().skt_coo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PlanesHandle(object):
    """Owner of the byte-digit planes of a dense device tensor (``skt_planes*``): the opt-in sliced contraction on the
    INT8 tensor cores (csrc/slice_gemm.cu).  ``nslices`` 6 = 48-bit, 5 = 40-bit fixed point under one scale."""

    def __init__(self, X_d, shape, nslices=6):
        lib = _lib.load()
        dev.require_cuda()
        self.shape = tuple(int(s) for s in shape)
        self.ndim = len(self.shape)
        self.nslices = int(nslices)
        h = C.c_void_p()
        _lib.check(lib.skt_planes_create(dev.ptr(X_d), _lib.shape_array(self.shape), self.ndim, self.nslices, C.byref(h),
                                         dev.stream_ptr()))
        self.h = h

    @staticmethod
    def supported(shape, mode, rank):
        shape = tuple(int(s) for s in shape)
        return bool(_lib.load().skt_planes_supported(_lib.shape_array(shape), len(shape), int(mode), int(rank)))

    def device_bytes(self):
        return int(_lib.load().skt_planes_device_bytes(self.h))

    def ttm(self, mode, U, rank, out=None):
        """``T[row, r] = sum_c X[.., c, ..] U[c, r]`` over the first or the last mode; rows in memory order."""
        lib = _lib.load()
        rows = 1
        for m, d in enumerate(self.shape):
            if m != mode:
                rows *= d
        if out is None:
            out = dev.empty((rows, rank))
        ws, wsz = dev.workspace('mttkrp').get(lib.skt_planes_ttm_workspace(self.h, mode, rank))
        _lib.check(lib.skt_planes_ttm(self.h, mode, dev.ptr(U), rank, dev.ptr(out), ws, wsz, dev.stream_ptr()))
        return out

    def fused_slabs(self, mode, rank):
        """(number of slabs, rows per slab) of ``ttm_fused`` for this mode, or (0, 0) when the shape is not supported."""
        rows = C.c_int64(0)
        n = int(_lib.load().skt_planes_fused_slabs(self.h, int(mode), int(rank), C.byref(rows)))
        return n, int(rows.value)

    def ttm_fused(self, mode, U, W, rank, slabs, out=None):
        """``ttm`` of a 3-way tensor with the tree's second stage folded in: ``slabs`` receives partial sums of the
        finished MTTKRP (mode 0 for ``mode = 2``, mode 2 for ``mode = 0``), ``W`` is the middle mode's factor."""
        lib = _lib.load()
        ws, wsz = dev.workspace('mttkrp').get(lib.skt_planes_ttm_workspace(self.h, mode, rank))
        _lib.check(lib.skt_planes_ttm_fused(self.h, mode, dev.ptr(U), dev.ptr(W), rank, dev.ptr(out) if out is not None else None,
                                            dev.ptr(slabs), slabs.numel() * slabs.element_size(), ws, wsz, dev.stream_ptr()))
        return out

    @staticmethod
    def pass_supported(shape, split, rank):
        shape = tuple(int(s) for s in shape)
        return bool(_lib.load().skt_planes_pass_supported(_lib.shape_array(shape), len(shape), int(split), int(rank)))

    def tree_pass(self, split, group, U, rank, out=None):
        """X as the matrix ``[prod(shape[:split]) x prod(shape[split:])]``: group 0 -> ``X KR(U[split:])``, group 1 ->
        ``X^T KR(U[:split])`` (Khatri-Rao product of the group's factors in memory order); rank <= 64, any length."""
        lib = _lib.load()
        P = 1
        for d in self.shape[:split]:
            P *= d
        Q = 1
        for d in self.shape[split:]:
            Q *= d
        if out is None:
            out = dev.empty((Q if group else P, rank))
        ws, wsz = dev.workspace('mttkrp').get(lib.skt_planes_pass_workspace(self.h, split, group, rank))
        _lib.check(lib.skt_planes_pass(self.h, split, group, _factor_ptrs(U), rank, dev.ptr(out), ws, wsz, dev.stream_ptr()))
        return out

    def close(self):
        if getattr(self, 'h', None):
            _lib.load().skt_planes_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_ACCUM_OPS = {'sum': 0, 'max': 1, 'min': 2}


def coo_accumulate(subs, vals, op):
    """Merge duplicate coordinates on the device (``utils.accum``, utils.py:5-26 with func = sum / max / min).

    Returns ``(vals, subs)`` with the distinct coordinates in the reference's order (lexsort: last mode most
    significant); values come back as float64 like the reference's ``np.zeros`` accumulator."""
    lib = _lib.load()
    dev.require_cuda()
    nd = len(subs)
    S = [np.ascontiguousarray(s, dtype=np.int64) for s in subs]
    v = np.ascontiguousarray(vals, dtype=np.float64)
    nnz = int(v.shape[0])
    base = [int(s.min()) for s in S]
    extent = [int(s.max()) - b + 1 for s, b in zip(S, base)]
    keys = np.empty(nnz, dtype=np.uint64)
    ovals = np.empty(nnz, dtype=np.float64)
    n_out = C.c_int64(0)
    ptrs = _lib.ptr_array([s.ctypes.data for s in S])
    _lib.check(lib.skt_coo_accumulate_host(ptrs, C.c_void_p(v.ctypes.data), nnz, _lib.shape_array(base),
                                           _lib.shape_array(extent), nd, _ACCUM_OPS[op], C.c_void_p(keys.ctypes.data),
                                           C.c_void_p(ovals.ctypes.data), C.byref(n_out)))
    m = n_out.value
    idx = np.unravel_index(keys[:m].astype(np.int64), tuple(extent), order='F')
    out_subs = tuple(np.asarray(ix, dtype=S[d].dtype) + base[d] for d, ix in enumerate(idx))
    return ovals[:m].copy(), out_subs


def compact_keys(keys_d, key_bits):
    """ids of ``keys_d`` (uint64 device tensor) among its distinct values -> (int32 device tensor, count)."""
    lib = _lib.load()
    n = int(keys_d.shape[0])
    ids = dev.empty((n,), np.int32)
    nd = C.c_int64(0)
    _lib.check(lib.skt_compact_keys(dev.ptr(keys_d), n, int(key_bits), dev.ptr(ids), C.byref(nd), dev.stream_ptr()))
    return ids, int(nd.value)


def mttkrp_coo(coo, U, rank, mode, out=None, atomic=False):
    """``sptensor.uttkrp`` on device factors (sptensor.py:216-229)."""
    lib = _lib.load()
    if out is None:
        out = dev.empty((coo.shape[mode], rank))
    ptrs = _factor_ptrs(U, mode)
    if atomic:
        _lib.check(lib.skt_mttkrp_coo_atomic(coo.h, ptrs, rank, mode, dev.ptr(out), dev.stream_ptr()))
        return out
    need = lib.skt_mttkrp_coo_workspace(coo.h, rank, mode)
    ws, wsz = dev.workspace('mttkrp').get(need)
    _lib.check(lib.skt_mttkrp_coo(coo.h, ptrs, rank, mode, dev.ptr(out), ws, wsz, dev.stream_ptr()))
    return out


def coo_innerprod(coo, U, lam, rank):
    lib = _lib.load()
    out = dev.empty((1,))
    ws, wsz = dev.workspace('reduce').get(lib.skt_coo_innerprod_workspace(coo.h, rank))
    _lib.check(lib.skt_coo_innerprod(coo.h, _factor_ptrs(U), dev.ptr(lam), rank, dev.ptr(out), ws, wsz,
                                     dev.stream_ptr()))
    return out


def coo_sumsq(coo):
    lib = _lib.load()
    out = dev.empty((1,))
    ws, wsz = dev.workspace('reduce').get(lib.skt_sumsq_workspace(coo.nnz))
    _lib.check(lib.skt_coo_sumsq(coo.h, dev.ptr(out), ws, wsz, dev.stream_ptr()))
    return out


# ------------------------------------------------------------------ reductions / normal equations
def sumsq(x):
    lib = _lib.load()
    out = dev.empty((1,))
    n = x.numel()
    ws, wsz = dev.workspace('reduce').get(lib.skt_sumsq_workspace(n))
    _lib.check(lib.skt_sumsq(dev.ptr(x), n, dev.ptr(out), ws, wsz, dev.stream_ptr()))
    return out


def weighted_dot(U, M, lam=None):
    """``sum_i sum_r lam_r U[i,r] M[i,r]`` -> 1-element device tensor."""
    lib = _lib.load()
    rows, R = U.shape
    out = dev.empty((1,))
    ws, wsz = dev.workspace('reduce').get(lib.skt_sumsq_workspace(rows * R))
    _lib.check(lib.skt_weighted_dot(dev.ptr(U), dev.ptr(M), dev.ptr(lam), rows, R, dev.ptr(out), ws, wsz,
                                    dev.stream_ptr()))
    return out


def gram(U, out=None):
    lib = _lib.load()
    rows, R = U.shape
    if out is None:
        out = dev.empty((R, R))
    ws, wsz = dev.workspace('pass').get(lib.skt_gram_workspace(rows, R))
    _lib.check(lib.skt_gram(dev.ptr(U), rows, R, dev.ptr(out), ws, wsz, dev.stream_ptr()))
    return out


def hadamard_pinv(G, mode, P=None, info=None, stream=None):
    """``pinv(hadamard_{m != mode} G[m])`` (cp.py:144-147); ``G`` is (ndim, R, R).  ``stream``: a torch
    stream to enqueue on instead of the current one."""
    lib = _lib.load()
    nd, R = int(G.shape[0]), int(G.shape[1])
    if P is None:
        P = dev.empty((R, R))
    if info is None:
        info = dev.empty((1,), np.int32)
    sp = dev.stream_ptr() if stream is None else C.c_void_p(stream.cuda_stream)
    _lib.check(lib.skt_hadamard_pinv(dev.ptr(G), nd, R, mode, dev.ptr(P), dev.ptr(info), sp))
    return P, info


def solve_stats(M, P, first_iter, Unew=None, colstat=None):
    lib = _lib.load()
    rows, R = M.shape
    if Unew is None:
        Unew = dev.empty((rows, R))
    if colstat is None:
        colstat = dev.empty((R,))
    ws, wsz = dev.workspace('pass').get(lib.skt_solve_stats_workspace(rows, R))
    _lib.check(lib.skt_solve_stats(dev.ptr(M), dev.ptr(P), rows, R, 1 if first_iter else 0, dev.ptr(Unew),
                                   dev.ptr(colstat), ws, wsz, dev.stream_ptr()))
    return Unew, colstat


def normalise_gram(U, colstat, first_iter, lam=None, G=None, Mip=None, ip=None):
    lib = _lib.load()
    rows, R = U.shape
    if lam is None:
        lam = dev.empty((R,))
    if G is None:
        G = dev.empty((R, R))
    if Mip is not None and ip is None:
        ip = dev.empty((1,))
    ws, wsz = dev.workspace('pass').get(lib.skt_normalise_gram_workspace(rows, R))
    _lib.check(lib.skt_normalise_gram(dev.ptr(U), dev.ptr(colstat), rows, R, 1 if first_iter else 0, dev.ptr(lam),
                                      dev.ptr(G), dev.ptr(Mip), dev.ptr(ip) if Mip is not None else None,
                                      ws, wsz, dev.stream_ptr()))
    return lam, G, ip


def cp_update_fused_fits(rows, rank):
    return bool(_lib.load().skt_cp_update_fused_fits(int(rows), int(rank)))


def cp_update_fused(M, G, mode, first_iter, U, lam, info, ip=None, normX2=None, fit=None):
    """pinv + solve + normalise + Gram (+ <P,X> and the fit) for a small factor in one launch (cp.py:144-159)."""
    lib = _lib.load()
    rows, R = M.shape
    _lib.check(lib.skt_cp_update_fused(dev.ptr(M), rows, R, int(G.shape[0]), mode, 1 if first_iter else 0, dev.ptr(G),
                                       dev.ptr(U), dev.ptr(lam), dev.ptr(info), dev.ptr(ip), dev.ptr(normX2),
                                       dev.ptr(fit), dev.stream_ptr()))


def cp_update_cluster_fits(rows, rank):
    return bool(_lib.load().skt_cp_update_cluster_fits(int(rows), int(rank)))


def cp_update_cluster(M, P, G, mode, first_iter, U, lam, ip=None, normX2=None, fit=None):
    """solve + normalise + Gram (+ <P,X>, fit) given P = pinv(Y), on one 8-CTA cluster (cp.py:147-159)."""
    lib = _lib.load()
    rows, R = M.shape
    _lib.check(lib.skt_cp_update_cluster(dev.ptr(M), dev.ptr(P), rows, R, int(G.shape[0]), mode, 1 if first_iter else 0,
                                         dev.ptr(G), dev.ptr(U), dev.ptr(lam), dev.ptr(ip), dev.ptr(normX2),
                                         dev.ptr(fit), dev.stream_ptr()))


def exchange_record(rank):
    return int(_lib.load().skt_cp_exchange_record(int(rank)))


def cp_update_exchange(M, P, first_iter, Unew, ex):
    """Phase A of the exchange-form update: ``Unew = M P`` and the record ``[colstat | Unew^T Unew | <Unew, M>]``."""
    lib = _lib.load()
    rows, R = M.shape
    ws, wsz = dev.workspace('pass').get(lib.skt_cp_update_exchange_workspace(rows, R))
    _lib.check(lib.skt_cp_update_exchange(dev.ptr(M), dev.ptr(P), rows, R, 1 if first_iter else 0, dev.ptr(Unew),
                                          dev.ptr(ex), ws, wsz, dev.stream_ptr()))


def cp_update_finish(ex_all, world, first_iter, U, lam, G, mode, ip=None, normX2=None, fit=None):
    """Phase B: lambda, ``G[mode]``, ``<P,X>``, fit from the gathered records; divides ``U`` by lambda if needed."""
    lib = _lib.load()
    rows, R = U.shape
    _lib.check(lib.skt_cp_update_finish(dev.ptr(ex_all), int(world), 1 if first_iter else 0, rows, R, int(G.shape[0]),
                                        mode, dev.ptr(U), dev.ptr(lam), dev.ptr(G), dev.ptr(ip), dev.ptr(normX2),
                                        dev.ptr(fit), dev.stream_ptr()))


def cp_update_cluster_parts(parts_dev, nparts, P, G, mode, first_iter, U, lam, ip=None, normX2=None, fit=None):
    """``cp_update_cluster`` with M = sum over ranks of the partial MTTKRPs read from peer memory (``parts_dev``: device
    address of the array of ``nparts`` peer pointers)."""
    lib = _lib.load()
    rows, R = U.shape
    _lib.check(lib.skt_cp_update_cluster_parts(C.c_void_p(int(parts_dev)), int(nparts), dev.ptr(P), rows, R, int(G.shape[0]),
                                               mode, 1 if first_iter else 0, dev.ptr(G), dev.ptr(U), dev.ptr(lam),
                                               dev.ptr(ip), dev.ptr(normX2), dev.ptr(fit), dev.stream_ptr()))


def cp_update_finish_parts(parts_dev, world, first_iter, U, lam, G, mode, ip=None, normX2=None, fit=None):
    """``cp_update_finish`` reading every rank's exchange record from peer memory."""
    lib = _lib.load()
    rows, R = U.shape
    _lib.check(lib.skt_cp_update_finish_parts(C.c_void_p(int(parts_dev)), int(world), 1 if first_iter else 0, rows, R,
                                              int(G.shape[0]), mode, dev.ptr(U), dev.ptr(lam), dev.ptr(G), dev.ptr(ip),
                                              dev.ptr(normX2), dev.ptr(fit), dev.stream_ptr()))


def cp_update_cluster_slabs(slabs, nslabs, P, G, mode, first_iter, U, lam, ip=None, normX2=None, fit=None):
    """``cp_update_cluster`` with M = sum of ``nslabs`` local partial slabs (``kr_reduce_slabs``)."""
    lib = _lib.load()
    rows, R = U.shape
    _lib.check(lib.skt_cp_update_cluster_slabs(dev.ptr(slabs), int(nslabs), dev.ptr(P), rows, R, int(G.shape[0]), mode,
                                               1 if first_iter else 0, dev.ptr(G), dev.ptr(U), dev.ptr(lam), dev.ptr(ip),
                                               dev.ptr(normX2), dev.ptr(fit), dev.stream_ptr()))


def set_reserved_sms(n):
    return int(_lib.load().skt_set_reserved_sms(int(n)))


def cp_fit(G, lam, ip, normX2, out=None):
    lib = _lib.load()
    if out is None:
        out = dev.empty((2,))
    _lib.check(lib.skt_cp_fit(dev.ptr(G), int(G.shape[0]), int(G.shape[1]), dev.ptr(lam), dev.ptr(ip),
                              dev.ptr(normX2), dev.ptr(out), dev.stream_ptr()))
    return out


# ------------------------------------------------------------------ HOOI pieces
def ttm_dense(Xd, shape, V, mode, transp):
    lib = _lib.load()
    p = int(V.shape[1] if transp else V.shape[0])
    oshape = list(shape)
    oshape[mode] = p
    Y = dev.empty(oshape)
    _lib.check(lib.skt_ttm_dense(dev.ptr(Xd), _lib.shape_array(shape), len(shape), dev.ptr(V), p, mode,
                                 1 if transp else 0, dev.ptr(Y), dev.stream_ptr()))
    return Y, tuple(oshape)


def ttm_dense_first_rot(Xd, shape, V):
    """``X x_0 V^T`` with the new mode stored last: shape ``(shape[1], ..., shape[-1], p)``."""
    lib = _lib.load()
    p = int(V.shape[1])
    oshape = tuple(int(s) for s in shape[1:]) + (p,)
    Y = dev.empty(oshape)
    _lib.check(lib.skt_ttm_dense_first_rot(dev.ptr(Xd), _lib.shape_array(shape), len(shape), dev.ptr(V), p,
                                           dev.ptr(Y), dev.stream_ptr()))
    return Y, oshape


def unfold_gram(Xd, shape, mode):
    lib = _lib.load()
    n = int(shape[mode])
    G = dev.empty((n, n))
    shp = _lib.shape_array(shape)
    ws, wsz = dev.workspace('reduce').get(lib.skt_unfold_gram_workspace(shp, len(shape), mode))
    _lib.check(lib.skt_unfold_gram(dev.ptr(Xd), shp, len(shape), mode, dev.ptr(G), ws, wsz, dev.stream_ptr()))
    return G


def cp_als_run_fits(shape, rank):
    return bool(_lib.load().skt_cp_als_run_fits(_lib.shape_array(shape), len(shape), int(rank)))


def cp_als_run(Xd, shape, rank, Ud, init_mask, max_iter, conv, fit_full):
    """The whole loop of cp.als in one C call (``skt_cp_als_run``) -> (lambda device tensor, fit, itr, exectimes)."""
    lib = _lib.load()
    shp = _lib.shape_array(shape)
    nd = len(shape)
    lam = dev.empty((rank,))
    ws, wsz = dev.workspace('run').get(lib.skt_cp_als_run_workspace(shp, nd, rank))
    fit, itr = C.c_double(0.0), C.c_int(-1)
    ex = np.zeros(max(int(max_iter), 1), dtype=np.float64)
    _lib.check(lib.skt_cp_als_run(dev.ptr(Xd), shp, nd, int(rank), _factor_ptrs(Ud), int(init_mask), int(max_iter),
                                  float(conv), 1 if fit_full else 0, dev.ptr(lam), ws, wsz, C.byref(fit), C.byref(itr),
                                  C.c_void_p(ex.ctypes.data), dev.stream_ptr()))
    return lam, fit.value, itr.value, ex[:itr.value + 1].copy()


SYMEIG_MAX_N = 2048


def symeig_top(G, k, flipsign=True, want_values=False):
    """Leading ``k`` eigenvectors of the symmetric device matrix ``G`` (n x n), descending, sign-fixed
    (core.py:284-306) -> device (n x k) [and the k eigenvalues].  One cluster kernel, no host synchronisation."""
    lib = _lib.load()
    n = int(G.shape[0])
    k = int(k)
    need = lib.skt_symeig_workspace(n, k)
    if need == 0:
        _lib.check(-6 if n > SYMEIG_MAX_N else -1)
    U = dev.empty((n, k))
    W = dev.empty((k,)) if want_values else None
    info = dev.empty((1,), np.int32)
    ws, wsz = dev.workspace('eig').get(need)
    _lib.check(lib.skt_symeig_top(dev.ptr(G), n, k, 1 if flipsign else 0, dev.ptr(W), dev.ptr(U), dev.ptr(info), ws, wsz,
                                  dev.stream_ptr()))
    return (U, W, info) if want_values else (U, info)


def probe_fp64_peaks():
    lib = _lib.load()
    dev.require_cuda()
    a, b = C.c_double(), C.c_double()
    _lib.check(lib.skt_probe_fp64_peaks(C.byref(a), C.byref(b)))
    return a.value, b.value

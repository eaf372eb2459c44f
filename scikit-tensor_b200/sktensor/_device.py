"""Device plumbing: memory, streams and workspaces (PyTorch is used for nothing else here).

The numerical work is done by ``libsktensor_b200.so`` through ``_lib``; this module only
moves NumPy arrays to HBM and back, hands out raw pointers, and keeps scratch buffers alive.
No function in here computes on the CPU: if CUDA is unavailable, ``require_cuda`` raises.
"""
import ctypes as C

import numpy as np

from . import _lib

_torch = None
_UPLOAD_MIN = 8 << 20


def torch():
    global _torch
    if _torch is None:
        import torch as _t
        _torch = _t
    return _torch


def require_cuda():
    t = torch()
    if not t.cuda.is_available():
        raise RuntimeError('sktensor (B200 build) needs a CUDA device: there is no CPU fallback')
    _lib.load()
    return t


def device():
    t = require_cuda()
    return t.device('cuda', t.cuda.current_device())


def stream_ptr():
    return C.c_void_p(torch().cuda.current_stream().cuda_stream)


def synchronize():
    torch().cuda.current_stream().synchronize()


def to_device(a, dtype=np.float64):
    """NumPy (any dtype / strides) -> contiguous C-order device tensor of ``dtype``."""
    t = require_cuda()
    if isinstance(a, t.Tensor):
        return a.to(device=device(), dtype=getattr(t, np.dtype(dtype).name)).contiguous()
    h = np.ascontiguousarray(np.asarray(a), dtype=dtype)
    if h.nbytes >= _UPLOAD_MIN:     # big tensors: multi-threaded pinned staging (skt_upload), PCIe-bound
        x = t.empty(h.shape, dtype=getattr(t, np.dtype(dtype).name), device=device())
        _lib.check(_lib.load().skt_upload(C.c_void_p(x.data_ptr()), C.c_void_p(h.ctypes.data), h.nbytes, stream_ptr()))
        return x
    if not h.flags.writeable:       # torch.from_numpy warns on read-only buffers
        h = h.copy()
    return t.from_numpy(h).to(device())


def to_host(x):
    return x.detach().cpu().numpy()


def empty(shape, dtype=np.float64):
    t = require_cuda()
    return t.empty(tuple(int(s) for s in shape) if not np.isscalar(shape) else (int(shape),),
                   dtype=getattr(t, np.dtype(dtype).name), device=device())


def zeros(shape, dtype=np.float64):
    x = empty(shape, dtype)
    x.zero_()
    return x


def ptr(x):
    """Raw device address of a tensor (``None`` -> NULL)."""
    return None if x is None else C.c_void_p(x.data_ptr())


class Workspace(object):
    """A grow-only scratch buffer; ``get(n)`` returns (pointer, size) of at least ``n`` bytes."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes):
        nbytes = max(int(nbytes), 256)
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = empty(nbytes + (nbytes >> 3), np.uint8)
        return C.c_void_p(self.buf.data_ptr()), C.c_size_t(self.buf.numel())


_ws = {}


def workspace(tag='default'):
    """Per-device named scratch buffers shared by the stateless entry points."""
    key = (torch().cuda.current_device(), tag)
    if key not in _ws:
        _ws[key] = Workspace()
    return _ws[key]

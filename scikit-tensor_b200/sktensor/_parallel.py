"""The CP-ALS sweep engine: device-resident state, kernel choreography, and mode sharding.

One process drives one GPU.  With ``torch.distributed`` initialised (world size > 1) the
tensor is cut along its largest mode ``s`` into contiguous index ranges -- equal rows for
dense tensors, equal non-zeros for sparse ones -- and each rank keeps its slab, its rows
of ``U_s`` and a replica of every other factor:

  * mode ``s``: the MTTKRP output rows are local; only the R-vector of column statistics
    and the R x R Gram of the updated rows are all-reduced;
  * mode ``n != s``: every rank produces a full-height partial MTTKRP from its slab, the
    partials are summed with one all-reduce, and the (small, replicated) solve is repeated
    on every rank, which leaves the replicas bitwise identical.

The collectives are issued through ``torch.distributed`` (NCCL on GPUs, gloo in the CPU
tests); nothing else in this file touches PyTorch.  The numerical kernels are reached
through a backend object so that the CPU test-suite can drive the same choreography with a
NumPy stand-in (``tests/fake_backend.py``); the shipped default is the CUDA library and
there is no automatic fallback.
"""
import os

import numpy as np

from .dtensor import dtensor
from .sptensor import sptensor


# ---------------------------------------------------------------------------------------------
# partitioning (pure host logic)
# ---------------------------------------------------------------------------------------------
def split_rows(n, world):
    """``world + 1`` offsets cutting ``range(n)`` into contiguous, nearly equal ranges."""
    base, extra = divmod(int(n), int(world))
    sizes = [base + (1 if r < extra else 0) for r in range(world)]
    return np.concatenate(([0], np.cumsum(sizes))).astype(np.int64)


def split_nnz(idx, dim, world, row_weight=None):
    """Offsets on ``range(dim)`` that balance ``nnz + row_weight * rows`` over the ranks.

    Power-law tensors put most non-zeros on few rows: equal-row cuts would leave one rank with nearly all the
    MTTKRP work, so the cuts follow the cumulative non-zero count.  Equal-nnz cuts alone leave the last rank with
    most of the ROWS, i.e. most of the factor update of the sharded mode (measured on config 3 at 8 GPUs: 3.3 M of
    10 M rows, 0.71 ms of a 4.3 ms iteration waiting for it), hence the row term: one row of the update costs about
    as much as two non-zeros of the three MTTKRPs (``SKT_SHARD_ROW_WEIGHT``, default 2).  Cuts never split a row.
    """
    if row_weight is None:
        row_weight = float(os.environ.get('SKT_SHARD_ROW_WEIGHT', '2'))
    counts = np.bincount(np.asarray(idx, dtype=np.int64), minlength=int(dim))
    cum = np.cumsum(counts).astype(np.float64) + row_weight * np.arange(1, int(dim) + 1, dtype=np.float64)
    total = float(cum[-1]) if len(cum) else 0.0
    offs = [0]
    for r in range(1, world):
        target = total * r / float(world)
        cut = int(np.searchsorted(cum, target, side='left')) + 1 if total else 0
        cut = min(max(cut, offs[-1]), int(dim))
        offs.append(cut)
    offs.append(int(dim))
    return np.array(offs, dtype=np.int64)


def shard_mode(shape, dense=False):
    """The mode to cut: the largest one.  Ties: the first -- except for dense tensors, where a middle mode wins: the
    first and the last mode are the GEMM contraction modes of the two dimension-tree passes, and a slab that is thin
    in a contraction mode leaves the tensor cores a K of a few rows per tile."""
    shape = [int(d) for d in shape]
    best = max(shape)
    cands = [m for m, d in enumerate(shape) if d == best]
    if dense:
        mid = [m for m in cands if 0 < m < len(shape) - 1]
        if mid:
            return mid[0]
    return cands[0]


def tree_split(shape):
    """Split point ``s`` of the two-group dimension tree: modes ``[0, s)`` share one partial
    contraction of X and modes ``[s, N)`` the other.  Chosen to keep both partial tensors small
    (minimal ``prod(shape[:s]) + prod(shape[s:])``; ties go to the larger ``s`` so that the second
    pass is the plain last-mode MTTKRP whose output also feeds the fit)."""
    N = len(shape)
    best, best_s = None, 1
    for s in range(1, N):
        cost = float(np.prod([float(d) for d in shape[:s]])) + float(np.prod([float(d) for d in shape[s:]]))
        if best is None or cost <= best:
            best, best_s = cost, s
    return best_s


class LocalShard(object):
    """A rank's slab of a tensor that is already partitioned along ``mode``.

    ``tensor`` is the local ``dtensor`` / ``sptensor`` (sparse subscripts along ``mode`` are
    LOCAL, i.e. already shifted by ``-offset``); ``global_shape`` is the full shape and
    ``offsets`` the ``world + 1`` cut points along ``mode``.
    """

    def __init__(self, tensor, mode, global_shape, offsets):
        self.tensor = tensor
        self.mode = int(mode)
        self.shape = tuple(int(s) for s in global_shape)
        self.ndim = len(self.shape)
        self.offsets = np.asarray(offsets, dtype=np.int64)


# ---------------------------------------------------------------------------------------------
# backends
# ---------------------------------------------------------------------------------------------
class CudaBackend(object):
    """The product backend: libsktensor_b200 kernels on the current CUDA device."""

    def __init__(self):
        from . import _device as dev
        from . import _kernels as K
        dev.require_cuda()
        self.dev, self.K = dev, K

    # memory
    def to_device(self, a): return self.dev.to_device(a)
    def to_host(self, x): return self.dev.to_host(x)
    def empty(self, shape, dtype=np.float64): return self.dev.empty(shape, dtype)
    def zeros(self, shape, dtype=np.float64): return self.dev.zeros(shape, dtype)
    # small result read-back that does not stall the stream: start now, wait later
    def async_read(self, tensors):
        t = self.dev.torch()
        host = [t.empty(x.shape, dtype=x.dtype, pin_memory=True) for x in tensors]
        for h, x in zip(host, tensors):
            h.copy_(x, non_blocking=True)
        ev = t.cuda.Event()
        ev.record()
        return ev, host

    def wait_read(self, handle):
        ev, host = handle
        ev.synchronize()
        return [h.numpy().copy() for h in host]

    # tensors
    def upload_dense(self, X):
        cached = getattr(X, '_dev', None)          # dtensor.cache_on_device(): reuse the resident copy
        return cached if cached is not None else self.dev.to_device(np.asarray(X))
    def upload_coo(self, subs, vals, shape): return self.K.CooHandle(subs, vals, shape)
    # kernels
    def mttkrp_dense(self, Xd, shape, U, R, n, out): return self.K.mttkrp_dense(Xd, shape, U, R, n, out=out)
    def mttkrp_coo(self, coo, U, R, n, out): return self.K.mttkrp_coo(coo, U, R, n, out=out)
    def kr_reduce(self, T, dims, W, R, n, out): return self.K.kr_reduce(T, dims, W, R, n, out=out)
    def planes_supported(self, shape, mode, R): return self.K.PlanesHandle.supported(shape, mode, R)
    def planes_create(self, Xd, shape, nslices): return self.K.PlanesHandle(Xd, shape, nslices)
    def planes_ttm(self, planes, mode, U, R, out): return planes.ttm(mode, U, R, out=out)
    def planes_fused_slabs(self, planes, mode, R): return planes.fused_slabs(mode, R)
    def planes_ttm_fused(self, planes, mode, U, W, R, slabs, out): return planes.ttm_fused(mode, U, W, R, slabs, out=out)
    def planes_pass_supported(self, shape, split, R): return self.K.PlanesHandle.pass_supported(shape, split, R)
    def planes_pass(self, planes, split, group, U, R, out): return planes.tree_pass(split, group, U, R, out=out)
    def kr_reduce_slab_buffer(self, dims, R, n): return self.empty((max(1, self.K.kr_reduce_workspace(dims, R, n) // 8),))
    def kr_reduce_slabs(self, T, dims, W, R, n, slabs): return self.K.kr_reduce_slabs(T, dims, W, R, n, slabs)
    def update_cluster_slabs(self, slabs, nslabs, P, G, n, first, U, lam, ip, normX2, fit):
        return self.K.cp_update_cluster_slabs(slabs, nslabs, P, G, n, first, U, lam, ip=ip, normX2=normX2, fit=fit)
    def sumsq(self, x): return self.K.sumsq(x)
    def coo_sumsq(self, coo): return self.K.coo_sumsq(coo)
    def gram(self, U, out): return self.K.gram(U, out=out)
    def hadamard_pinv(self, G, n, P, info): return self.K.hadamard_pinv(G, n, P=P, info=info)
    def solve_stats(self, M, P, first, Unew, colstat): return self.K.solve_stats(M, P, first, Unew=Unew, colstat=colstat)
    def normalise_gram(self, U, colstat, first, lam, G, Mip, ip):
        return self.K.normalise_gram(U, colstat, first, lam=lam, G=G, Mip=Mip, ip=ip)
    def cp_fit(self, G, lam, ip, normX2, out): return self.K.cp_fit(G, lam, ip, normX2, out=out)
    # the short update on an 8-CTA cluster, with pinv(Y) running on a side stream during the MTTKRP
    def update_cluster_fits(self, rows, R): return self.K.cp_update_cluster_fits(rows, R)

    def pinv_async(self, G, n, P, info):
        t = self.dev.torch()
        if getattr(self, '_side', None) is None:
            self._side = t.cuda.Stream()
            self._ev = {}
            self.K.set_reserved_sms(1)      # the persistent MTTKRP grid leaves one SM to the side stream
        ready, done = self._ev.setdefault(n, (t.cuda.Event(), t.cuda.Event()))
        ready.record()                      # the Grams this pinv reads are final on the main stream
        self._side.wait_event(ready)
        self.K.hadamard_pinv(G, n, P=P, info=info, stream=self._side)
        done.record(self._side)
        return done

    def wait_pinv(self, handle):
        self.dev.torch().cuda.current_stream().wait_event(handle)

    def release_side(self):
        if getattr(self, '_side', None) is not None:
            self.dev.torch().cuda.current_stream().wait_stream(self._side)
            self.K.set_reserved_sms(0)
            self._side = None

    def update_cluster(self, M, P, G, n, first, U, lam, ip, normX2, fit):
        return self.K.cp_update_cluster(M, P, G, n, first, U, lam, ip=ip, normX2=normX2, fit=fit)

    # exchange-form update (sharded / tall factors): one record per rank, one collective, one finishing kernel
    def exchange_record(self, R): return self.K.exchange_record(R)
    def update_exchange(self, M, P, first, Unew, ex): return self.K.cp_update_exchange(M, P, first, Unew, ex)
    def update_finish(self, ex_all, world, first, U, lam, G, n, ip, normX2, fit):
        return self.K.cp_update_finish(ex_all, world, first, U, lam, G, n, ip=ip, normX2=normX2, fit=fit)

    # symmetric memory (NVLink peer access): buffers every rank can read inside a kernel
    _SYMM = {}      # (device, shape, tag) -> (tensor, handle): the rendezvous costs ~0.1 s, so buffers outlive the engines

    def symm_alloc(self, shape, tag=0):
        """(tensor, handle) in symmetric memory, or None when the platform lacks it.  Collective: all ranks call it
        with the same arguments in the same order.  Buffers are cached per process and reused by later calls."""
        import torch.distributed as dist
        mode = os.environ.get('SKT_SYMM', '')            # '1' force, '0' off; default: from 4 ranks up, where the NCCL
        if mode == '0' or os.environ.get('SKT_NO_SYMM', '0') == '1':   # latency it replaces dominates (measured: 2 GPUs
            return None                                  # 0.752 vs 0.743 ms/step, 8 GPUs 0.765 vs 0.848)
        if mode != '1' and dist.get_world_size() < 4:
            return None
        key = (self.dev.torch().cuda.current_device(), tuple(int(d) for d in shape), tag)
        if key in CudaBackend._SYMM:
            return CudaBackend._SYMM[key]      # (the outcome, None included, was agreed on by all ranks below)
        t = self.dev.torch()
        got = None
        try:
            import torch.distributed._symmetric_memory as symm_mem
            buf = symm_mem.empty(key[1], dtype=t.float64, device=self.dev.device())
            buf.zero_()
            hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
            got = (buf, hdl)
        except Exception:
            got = None
        # a rank that failed alone would take the NCCL path while the others wait in the peer-memory barrier:
        # every rank uses the peer-memory path only if every rank got its buffer
        ok = t.tensor([1 if got is not None else 0], dtype=t.int32, device=self.dev.device())
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            got = None
        CudaBackend._SYMM[key] = got
        return got

    def symm_barrier(self, hdl, channel=0):
        hdl.barrier(channel=channel)            # cross-rank barrier on the current stream

    def update_cluster_parts(self, hdl, world, P, G, n, first, U, lam, ip, normX2, fit):
        return self.K.cp_update_cluster_parts(hdl.buffer_ptrs_dev, world, P, G, n, first, U, lam, ip=ip, normX2=normX2, fit=fit)

    def update_finish_parts(self, hdl, world, first, U, lam, G, n, ip, normX2, fit):
        return self.K.cp_update_finish_parts(hdl.buffer_ptrs_dev, world, first, U, lam, G, n, ip=ip, normX2=normX2, fit=fit)

    def update_fused_fits(self, rows, R): return self.K.cp_update_fused_fits(rows, R)
    def update_fused(self, M, G, n, first, U, lam, info, ip, normX2, fit):
        return self.K.cp_update_fused(M, G, n, first, U, lam, info, ip=ip, normX2=normX2, fit=fit)


_backend_factory = CudaBackend


def set_backend_factory(factory):
    """Replace the kernel backend constructor (test hook; the default is ``CudaBackend``)."""
    global _backend_factory
    _backend_factory = factory if factory is not None else CudaBackend


class _Comm(object):
    """torch.distributed wrapper; a no-op object when running on a single rank."""

    def __init__(self):
        self.rank, self.world, self.dist = 0, 1, None
        try:
            import torch.distributed as dist
        except Exception:      # torch missing is caught later by the backend
            return
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            self.dist = dist
            self.rank = dist.get_rank()
            self.world = dist.get_world_size()

    @property
    def active(self):
        return self.dist is not None

    def allreduce_sum(self, x):
        if self.active:
            self.dist.all_reduce(x, op=self.dist.ReduceOp.SUM)

    def allreduce_max(self, x):
        if self.active:
            self.dist.all_reduce(x, op=self.dist.ReduceOp.MAX)

    def allgather_records(self, rec, out):
        """``out[w] = rec`` of rank ``w`` (equal-sized records); a copy on a single rank."""
        if not self.active:
            out[0].copy_(rec)
            return
        try:
            self.dist.all_gather_into_tensor(out.view(-1), rec)
        except Exception:           # backends without the flat variant (older gloo)
            parts = [out[w] for w in range(self.world)]
            self.dist.all_gather(parts, rec)

    def reduce_scatter_rows(self, full, out):
        """``out`` (chunk x R) = this rank's chunk of the sum over ranks of ``full`` (world * chunk x R)."""
        try:
            self.dist.reduce_scatter_tensor(out, full, op=self.dist.ReduceOp.SUM)
        except Exception:           # backends without reduce-scatter (gloo): all-reduce, keep the own chunk
            self.dist.all_reduce(full, op=self.dist.ReduceOp.SUM)
            c = out.shape[0]
            out.copy_(full[self.rank * c:(self.rank + 1) * c])

    def allgather_chunks(self, chunk, full):
        """``full`` (world * chunk x R) = the ranks' ``chunk`` buffers, in rank order."""
        try:
            self.dist.all_gather_into_tensor(full, chunk)
        except Exception:
            c = chunk.shape[0]
            self.dist.all_gather([full[w * c:(w + 1) * c] for w in range(self.world)], chunk)

    def allgather_rows(self, local, offsets, backend):
        """Concatenate the ranks' row blocks (sizes from ``offsets``) into one host array."""
        if not self.active:
            return backend.to_host(local)
        sizes = np.diff(offsets)
        maxr = int(sizes.max())
        pad = backend.zeros((maxr, local.shape[1]))
        pad[:local.shape[0]] = local
        bufs = [backend.empty((maxr, local.shape[1])) for _ in range(self.world)]
        self.dist.all_gather(bufs, pad)
        return np.concatenate([backend.to_host(b)[:int(sizes[r])] for r, b in enumerate(bufs)], axis=0)


# ---------------------------------------------------------------------------------------------
# the engine
# ---------------------------------------------------------------------------------------------
class CpAlsEngine(object):
    """Device-resident CP-ALS state; ``sweep`` runs one iteration of cp.py:142-159."""

    def __init__(self, X, rank, U, backend=None, comm=None):
        self.be = backend if backend is not None else _backend_factory()
        self.comm = comm if comm is not None else _Comm()
        be, comm = self.be, self.comm
        self.R = R = int(rank)
        if R < 1:
            raise ValueError('rank must be a positive integer')
        if R > 512:
            raise ValueError('rank %d is beyond the 512 components the solve kernels of this build support' % R)

        # ---- which slab is ours
        if isinstance(X, LocalShard):
            self.shape, self.s, self.offsets = X.shape, X.mode, X.offsets
            local = X.tensor
            if not comm.active:          # a "shard" that is the whole tensor: nothing is sharded
                self.s = -1
        else:
            self.shape = tuple(int(d) for d in X.shape)
            self.s = shard_mode(self.shape, dense=isinstance(X, dtensor)) if comm.active else -1
            local = X
            if comm.active:
                local, self.offsets = self._cut(X, self.s, comm)
            else:
                self.offsets = None
        self.N = N = len(self.shape)
        self.sparse = isinstance(local, sptensor)
        if not self.sparse and not isinstance(local, dtensor):
            raise ValueError('cp_als needs a dtensor or sptensor (got %s)' % type(local))
        self.lshape = tuple(int(d) for d in local.shape)
        lo = int(self.offsets[comm.rank]) if comm.active else 0
        hi = int(self.offsets[comm.rank + 1]) if comm.active else 0
        self.row_lo, self.row_hi = lo, hi

        # ---- tensor to the device, ||X||^2
        if self.sparse:
            cached = getattr(local, '_coo', None)        # sptensor keeps its sorted device copies once built
            self.Xd = cached if cached is not None else \
                be.upload_coo([np.asarray(s_) for s_ in local.subs], local.vals, self.lshape)
            nnz = getattr(self.Xd, 'nnz', len(local.vals))
            self.normX2 = be.coo_sumsq(self.Xd) if nnz else be.zeros((1,))
        else:
            self.Xd = be.upload_dense(local)
            self.normX2 = be.sumsq(self.Xd)
        comm.allreduce_sum(self.normX2)

        # ---- factors (local rows for the sharded mode), Grams, scratch
        self.Ud = [None] * N
        self.G = be.zeros((N, R, R))
        for m in range(N):
            if U[m] is None:
                continue
            Um = np.asarray(U[m], dtype=np.float64)
            if Um.ndim != 2 or Um.shape != (self.shape[m], R):
                raise ValueError('Dimension mismatch of factor matrices')
            if m == self.s:
                Um = Um[lo:hi]
            self.Ud[m] = be.to_device(Um)
            be.gram(self.Ud[m], self.G[m])
            if m == self.s:
                comm.allreduce_sum(self.G[m])
        for m in range(N):
            if self.Ud[m] is None:
                self.Ud[m] = be.zeros((self.lshape[m], R))
        self.missing = [m for m in range(N) if U[m] is None]
        self.M = [be.empty((self.lshape[m], R)) for m in range(N)]
        self.P = be.empty((R, R))
        self.info = be.zeros((N,), np.int32)
        self.colstat = be.empty((R,))
        self.lam = be.empty((R,))
        self.ip = be.zeros((1,))
        self.fitout = be.zeros((2,))
        self._m0_ready = False      # True when M[0] already holds the MTTKRP for the coming sweep
        self._pending = None

        # ---- modes whose whole update runs in the fused single-launch kernel: the factor fits in shared
        # memory and the solve is replicated (the sharded mode needs collectives between the phases)
        use_fused = hasattr(be, 'update_fused') and os.environ.get('SKT_NO_FUSED_UPDATE', '0') != '1'
        self.fused = [bool(use_fused and not (comm.active and m == self.s) and be.update_fused_fits(self.lshape[m], R))
                      for m in range(N)]

        # ---- ... and, preferably, on an 8-CTA cluster with pinv(Y_n) computed on a side stream while the MTTKRP
        # of mode n is still streaming the tensor (Y_n only needs the Grams of the other modes)
        use_cluster = use_fused and hasattr(be, 'update_cluster') and os.environ.get('SKT_NO_CLUSTER_UPDATE', '0') != '1'
        self.cluster = [bool(use_cluster and self.fused[m] and be.update_cluster_fits(self.lshape[m], R)) for m in range(N)]
        self._pinv_pending = [None] * N

        # ---- every other mode (sharded, or too tall for shared memory) takes the exchange form when available; ranks
        # beyond the single-CTA solve kernels (> 64) run the generic gram / pinv / solve / normalise entry points instead
        self.exchange = hasattr(be, 'update_exchange') and os.environ.get('SKT_NO_EXCHANGE_UPDATE', '0') != '1' and \
            R <= getattr(be, 'max_exchange_rank', 64)
        if self.exchange:
            rec = be.exchange_record(R)
            self.ex = be.zeros((rec,))
            self.ex_all = be.zeros((comm.world if comm.active else 1, rec))
        # ---- multi-GPU: fuse the reductions into the update kernels over NVLink peer memory.  The partial MTTKRP of a
        # small replicated-solve mode and the exchange record of the sharded mode live in symmetric memory; after a
        # cross-rank stream barrier the update kernels add the peers' buffers themselves, in rank order (no NCCL call).
        self.Mh = [None] * N
        self.exh = None
        if comm.active and hasattr(be, 'symm_alloc'):
            ok = True
            for m in range(N):
                if ok and m != self.s and self.cluster[m]:
                    got = be.symm_alloc((self.lshape[m], R), tag=m)
                    if got is None:
                        ok = False
                    else:
                        self.M[m], self.Mh[m] = got
            if ok and self.exchange and not self.fused[self.s]:
                got = be.symm_alloc((be.exchange_record(R),), tag='ex')
                if got is not None:
                    self.ex, self.exh = got

        # ---- multi-GPU, tall replicated factor (sparse configs: 10^5 .. 10^6 rows): instead of all-reducing the full-height
        # partial M_n and repeating the update of every row on every rank, reduce-scatter M_n, update the scattered rows in
        # exchange form (one record per rank), all-gather the updated rows.  Same bytes on the wire as the
        # all-reduce; the update work and its HBM traffic drop by the world size.
        self.rs = [False] * N
        self.rs_chunk = [0] * N
        self.Mpad = [None] * N
        self.Mloc = [None] * N
        self.Upad = [None] * N
        self.Uloc = [None] * N
        if comm.active and self.exchange and os.environ.get('SKT_NO_SCATTERED_UPDATE', '0') != '1':
            for m in range(N):
                if m == self.s or self.cluster[m] or self.fused[m] or self.Mh[m] is not None:
                    continue
                rows = self.lshape[m]
                if rows < int(os.environ.get('SKT_SCATTER_MIN_ROWS', '32768')):
                    continue
                chunk = (rows + comm.world - 1) // comm.world
                self.rs[m], self.rs_chunk[m] = True, chunk
                self.Mpad[m] = be.zeros((chunk * comm.world, R))
                self.M[m] = self.Mpad[m][:rows]                 # the MTTKRP writes (and clears) the first `rows` rows only
                self.Mloc[m] = be.zeros((chunk, R))
                self.Upad[m] = be.zeros((chunk * comm.world, R))
                self.Upad[m][:rows].copy_(self.Ud[m])
                self.Ud[m] = self.Upad[m][:rows]
                self.Uloc[m] = be.zeros((chunk, R))

        # modes whose pinv runs on the side stream: every mode that consumes a ready-made P
        has_async = hasattr(be, 'pinv_async') and os.environ.get('SKT_NO_CLUSTER_UPDATE', '0') != '1'
        self.apinv = [bool(has_async and (self.cluster[m] or (self.exchange and not self.fused[m]))) for m in range(N)]
        self.Pn = [be.empty((R, R)) if self.apinv[m] else None for m in range(N)]

        # ---- dimension tree (dense, N >= 3): two passes over X per sweep instead of N.  Modes [0, s) are
        # served by T_A = X contracted with U_s..U_{N-1}; modes [s, N) by T_B = X contracted with
        # U_0..U_{s-1}.  A one-mode group needs no second stage: its partial tensor IS the MTTKRP.
        self.slabbuf = [None] * N        # per mode: buffer for un-summed second-stage slabs (see below)
        self.i8 = None                   # byte planes of X for the opt-in sliced passes
        self.i8_general = False
        self.i8_fused = [False] * N      # modes whose MTTKRP arrives as slabs from a fused sliced pass
        self.nslabs = [0] * N
        self.tree = (not self.sparse) and N >= 3 and hasattr(be, 'kr_reduce') and \
            os.environ.get('SKT_NO_DIMTREE', '0') != '1'
        if self.tree:
            self.ts = s = tree_split(self.lshape)
            nA = int(np.prod(self.lshape[:s]))
            nB = int(np.prod(self.lshape[s:]))
            self.shapeA = (nA,) + tuple(self.lshape[s:])
            self.shapeB = tuple(self.lshape[:s]) + (nB,)
            self.TA = self.M[0] if s == 1 else be.empty((nA, R))
            self.TB = self.M[N - 1] if s == N - 1 else be.empty((nB, R))
            # opt-in (SKT_DENSE_I8=6 | 5 | 1): both passes of a 3-way tree as sliced contractions on the INT8 tensor cores
            # (csrc/slice_gemm.cu).  X is cut into byte planes once; a pass whose group holds two modes becomes the GEMM
            # over the outer mode followed by the second-stage reduction the tree already has.
            # Other shapes (4-way and up, rank 33..64, long modes): X as the matrix [prod(shape[:s]) x prod(shape[s:])], each
            # pass one long GEMM against the Khatri-Rao product of the group's factors (skt_planes_pass).
            want = os.environ.get('SKT_DENSE_I8', '0')
            self.i8_general = False
            if want not in ('', '0') and hasattr(be, 'planes_create'):
                if N == 3 and be.planes_supported(self.lshape, 0, R) and be.planes_supported(self.lshape, N - 1, R):
                    self.i8 = be.planes_create(self.Xd, self.lshape, 5 if want == '5' else 6)
                    # single process: the second stage of the last mode rides in the epilogue of the first-mode contraction
                    # and arrives as slabs that the cluster update sums while it stages M (no partial tensor, no kr_reduce)
                    if s == N - 1 and not comm.active and hasattr(be, 'planes_ttm_fused') and \
                            os.environ.get('SKT_I8_NO_FUSED', '0') != '1':
                        ns, rows = be.planes_fused_slabs(self.i8, 0, R)
                        if ns > 0 and self.cluster[N - 1] and rows == self.lshape[N - 1]:
                            self.i8_fused[N - 1] = True
                            self.slabbuf[N - 1] = be.empty((ns, rows, R))
                            self.nslabs[N - 1] = ns
                    if not self.i8_fused[N - 1]:                # partial tensor of the two-mode group's GEMM
                        keep = self.lshape[1:] if s == N - 1 else self.lshape[:N - 1]
                        self.T2 = be.empty((int(np.prod(keep)), R))
                elif be.planes_pass_supported(self.lshape, s, R):
                    self.i8 = be.planes_create(self.Xd, self.lshape, 5 if want == '5' else 6)
                    self.i8_general = True
            # single process: the second stage leaves its per-split slabs for the cluster update to add while it stages
            # M (same arithmetic, one launch less on the critical path)
            if not comm.active and hasattr(be, 'kr_reduce_slabs') and os.environ.get('SKT_NO_KR_SLABS', '0') != '1':
                for m in range(N):
                    two_stage = (m < s and s > 1) or (m >= s and s < N - 1)
                    if two_stage and self.cluster[m] and not self.i8_fused[m]:
                        dims = self.lshape[:s] if m < s else self.lshape[s:]
                        self.slabbuf[m] = be.kr_reduce_slab_buffer(dims, R, m if m < s else m - s)

    @staticmethod
    def _cut(X, s, comm):
        """Slice a full tensor down to this rank's slab along mode ``s``."""
        if isinstance(X, sptensor):
            offs = split_nnz(X.subs[s], X.shape[s], comm.world)
            lo, hi = int(offs[comm.rank]), int(offs[comm.rank + 1])
            idx = np.asarray(X.subs[s])
            keep = (idx >= lo) & (idx < hi)
            subs = tuple((np.asarray(sm)[keep] - (lo if m == s else 0)).astype(np.int64)
                         for m, sm in enumerate(X.subs))
            shape = list(X.shape)
            shape[s] = hi - lo
            return sptensor(subs, np.asarray(X.vals)[keep], shape=shape), offs
        offs = split_rows(X.shape[s], comm.world)
        lo, hi = int(offs[comm.rank]), int(offs[comm.rank + 1])
        sl = [slice(None)] * X.ndim
        sl[s] = slice(lo, hi)
        return dtensor(np.ascontiguousarray(np.asarray(X)[tuple(sl)])), offs

    # ------------------------------------------------------------------ one ALS iteration
    def _tree_pass(self, group):
        """One pass over X: the partial contraction that serves every mode of ``group`` (0: modes
        ``[0, s)``, 1: modes ``[s, N)``), by the dense MTTKRP kernel on a reshaped view of X."""
        be, s, N = self.be, self.ts, self.N
        if self.i8 is not None and self.i8_general:
            be.planes_pass(self.i8, s, group, self.Ud, self.R, self.TA if group == 0 else self.TB)
            return
        if self.i8 is not None:
            R, Ud, ls = self.R, self.Ud, self.lshape
            if group == 1 and s == N - 1 and self.i8_fused[N - 1]:
                be.planes_ttm_fused(self.i8, 0, Ud[0], Ud[1], R, self.slabbuf[N - 1], None)
            elif group == 0 and s == N - 1:
                be.planes_ttm(self.i8, N - 1, Ud[N - 1], R, self.TA)
            elif group == 0:                                    # s == 1: T_A = M_0
                be.planes_ttm(self.i8, N - 1, Ud[N - 1], R, self.T2)
                be.kr_reduce(self.T2, ls[:N - 1], Ud[:N - 1], R, 0, self.TA)
            elif s == 1:
                be.planes_ttm(self.i8, 0, Ud[0], R, self.TB)
            else:                                               # s == N - 1: T_B = M_{N-1}
                be.planes_ttm(self.i8, 0, Ud[0], R, self.T2)
                be.kr_reduce(self.T2, ls[1:], Ud[1:], R, N - 2, self.TB)
            return
        if group == 0:
            be.mttkrp_dense(self.Xd, self.shapeA, [None] + list(self.Ud[s:]), self.R, 0, self.TA)
        else:
            be.mttkrp_dense(self.Xd, self.shapeB, list(self.Ud[:s]) + [None], self.R, s, self.TB)

    def mttkrp(self, n):
        """Local MTTKRP for mode ``n`` into ``self.M[n]`` (partial sums reduced across ranks)."""
        be = self.be
        # the reference ignores U[n]; on the first sweep the not-yet-computed factors are never read
        if self.sparse:
            be.mttkrp_coo(self.Xd, self.Ud, self.R, n, self.M[n])
        elif self.tree:
            s, N = self.ts, self.N
            if n == 0:
                self._tree_pass(0)
            elif n == s:
                self._tree_pass(1)
            if self.i8_fused[n]:
                pass                                            # the pass left the slabs of this mode
            elif n < s and s > 1:
                if self.slabbuf[n] is not None:
                    self.nslabs[n] = be.kr_reduce_slabs(self.TA, self.lshape[:s], self.Ud[:s], self.R, n, self.slabbuf[n])
                else:
                    be.kr_reduce(self.TA, self.lshape[:s], self.Ud[:s], self.R, n, self.M[n])
            elif n >= s and s < N - 1:
                if self.slabbuf[n] is not None:
                    self.nslabs[n] = be.kr_reduce_slabs(self.TB, self.lshape[s:], self.Ud[s:], self.R, n - s, self.slabbuf[n])
                else:
                    be.kr_reduce(self.TB, self.lshape[s:], self.Ud[s:], self.R, n - s, self.M[n])
        else:
            be.mttkrp_dense(self.Xd, self.lshape, self.Ud, self.R, n, self.M[n])
        if self.comm.active and n != self.s and self.Mh[n] is None and not self.rs[n]:
            self.comm.allreduce_sum(self.M[n])          # (modes with a symmetric-memory M are summed inside the update kernel;
                                                        #  tall modes are reduce-scattered in sweep())
        return self.M[n]

    def _mark(self, label):
        """SKT_SWEEP_TRACE=1: CUDA events at the phase boundaries of a sweep (see trace_report)."""
        tr = getattr(self, '_trace', None)
        if tr is None:
            return
        ev = self.be.dev.torch().cuda.Event(enable_timing=True)
        ev.record()
        tr.append((label, ev))

    def trace_begin(self):
        self._trace = []

    def trace_report(self):
        """ms per phase since trace_begin(), summed over the sweeps traced (call after a synchronise)."""
        tr, self._trace = self._trace, None
        out = {}
        for (l0, e0), (l1, e1) in zip(tr[:-1], tr[1:]):
            out[l1] = out.get(l1, 0.0) + e0.elapsed_time(e1)
        return out

    def sweep(self, first_iter, want_fit=True):
        be, comm, N = self.be, self.comm, self.N
        fit_done = False
        self._mark('start')

        def start_pinv(m):
            # pinv(Y_m) only needs the Grams of the other modes: run it on the side stream beside the MTTKRP of mode m
            if self.apinv[m] and self._pinv_pending[m] is None:
                self._pinv_pending[m] = be.pinv_async(self.G, m, self.Pn[m], self.info[m:m + 1])

        for n in range(N):
            start_pinv(n)                                   # (only the first mode of the first sweep starts here)
            if n == 0 and self._m0_ready:
                M, self._m0_ready = self.M[0], False
            else:
                M = self.mttkrp(n)
            self._mark('mttkrp%d' % n)
            last = (n == N - 1) and want_fit
            fit_args = (self.ip if last else None, self.normX2 if last else None, self.fitout if last else None)
            if self.apinv[n]:
                be.wait_pinv(self._pinv_pending[n])
                self._pinv_pending[n] = None
                P = self.Pn[n]
            elif self.cluster[n] or not self.fused[n]:
                be.hadamard_pinv(self.G, n, self.P, self.info[n:n + 1])
                P = self.P
            if self.cluster[n]:
                # small factor, replicated solve: one 8-CTA cluster does solve + normalise + Gram (+ <P,X>, fit)
                if self.slabbuf[n] is not None:
                    be.update_cluster_slabs(self.slabbuf[n], self.nslabs[n], P, self.G, n, first_iter, self.Ud[n], self.lam,
                                            *fit_args)
                elif self.Mh[n] is not None:
                    be.symm_barrier(self.Mh[n])             # every rank's partial M_n is complete and visible
                    be.update_cluster_parts(self.Mh[n], comm.world, P, self.G, n, first_iter, self.Ud[n], self.lam, *fit_args)
                else:
                    be.update_cluster(M, P, self.G, n, first_iter, self.Ud[n], self.lam, *fit_args)
                fit_done = fit_done or last
            elif self.fused[n]:
                # the same in one CTA, pinv included (no side stream)
                be.update_fused(M, self.G, n, first_iter, self.Ud[n], self.lam, self.info[n:n + 1], *fit_args)
                fit_done = fit_done or last
            elif self.exchange and self.rs[n]:
                # tall replicated factor on several GPUs: scattered update (see __init__)
                chunk, rows = self.rs_chunk[n], self.lshape[n]
                nloc = max(0, min(chunk, rows - comm.rank * chunk))
                comm.reduce_scatter_rows(self.Mpad[n], self.Mloc[n])
                be.update_exchange(self.Mloc[n][:nloc], P, first_iter, self.Uloc[n][:nloc], self.ex)
                comm.allgather_records(self.ex, self.ex_all)
                be.update_finish(self.ex_all, comm.world, first_iter, self.Uloc[n][:nloc], self.lam, self.G, n, *fit_args)
                comm.allgather_chunks(self.Uloc[n], self.Upad[n])
                fit_done = fit_done or last
            elif self.exchange:
                # sharded mode / tall factor: one record per rank, one all-gather, one finishing kernel
                sharded = comm.active and n == self.s
                be.update_exchange(M, P, first_iter, self.Ud[n], self.ex if sharded else self.ex_all[0])
                if sharded and self.exh is not None:
                    be.symm_barrier(self.exh)
                    be.update_finish_parts(self.exh, comm.world, first_iter, self.Ud[n], self.lam, self.G, n, *fit_args)
                else:
                    if sharded:
                        comm.allgather_records(self.ex, self.ex_all)
                    be.update_finish(self.ex_all, comm.world if sharded else 1, first_iter, self.Ud[n], self.lam, self.G, n,
                                     *fit_args)
                fit_done = fit_done or last
            else:
                be.solve_stats(M, P, first_iter, self.Ud[n], self.colstat)
                if comm.active and n == self.s:
                    if first_iter:
                        comm.allreduce_sum(self.colstat)
                    else:
                        comm.allreduce_max(self.colstat)
                be.normalise_gram(self.Ud[n], self.colstat, first_iter, self.lam, self.G[n],
                                  M if last else None, self.ip if last else None)
                if comm.active and n == self.s:
                    comm.allreduce_sum(self.G[n])
                    if last:
                        comm.allreduce_sum(self.ip)
            self._mark('update%d' % n)
            start_pinv((n + 1) % N)                         # the Grams it needs are final now
        if want_fit and not fit_done:
            be.cp_fit(self.G, self.lam, self.ip, self.normX2, self.fitout)

    def begin_read(self):
        """Start copying (fit, status) to the host without stalling the stream."""
        if hasattr(self.be, 'async_read'):
            self._pending = self.be.async_read([self.fitout, self.info])

    def prefetch_next_sweep(self):
        """Enqueue the next sweep's mode-0 MTTKRP now (it reads the current factors and writes only
        M[0]), so the GPU has work while the host waits for the fit; harmless if no sweep follows."""
        self._mark('start')
        self.mttkrp(0)
        self._mark('mttkrp0(prefetched)')
        self._m0_ready = True

    def _results(self):
        if self._pending is not None:
            fitout, info = self.be.wait_read(self._pending)
            self._pending = None
            return fitout, info
        return self.be.to_host(self.fitout), self.be.to_host(self.info)

    def check_status(self, info=None):
        if info is None:
            _, info = self._results()
        if (info < 0).any():
            # scipy.linalg.pinv raises exactly this on a non-finite Gram product (cp.py:147)
            raise ValueError('array must not contain infs or NaNs')

    def read_fit(self):
        fitout, info = self._results()
        self.check_status(info)
        return float(fitout[0])

    def release(self):
        """Give back what the engine holds process-wide (the SM reserved for the side-stream pinv)."""
        if hasattr(self.be, 'release_side'):
            self.be.release_side()          # a speculative pinv of the next sweep may still be running
        self._pinv_pending = [None] * self.N
        if getattr(self, 'i8', None) is not None:
            self.i8.close()
            self.i8 = None

    def download(self, U):
        """Write the factors into the caller's list (in place) and return lambda."""
        self.release()
        for m in range(self.N):
            if self.comm.active and m == self.s:
                U[m] = self.comm.allgather_rows(self.Ud[m], self.offsets, self.be)
            else:
                U[m] = self.be.to_host(self.Ud[m])
        return self.be.to_host(self.lam)

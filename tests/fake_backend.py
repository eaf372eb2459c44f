"""NumPy stand-in for the CUDA kernel backend -- TEST DOUBLE ONLY.

It lets the CPU test-suite drive ``sktensor._parallel.CpAlsEngine`` (sharding, collective
placement, status handling) under ``gloo`` with world_size > 1, where no GPU exists.  Storage
is CPU ``torch`` tensors so ``torch.distributed`` can reduce them; arithmetic comes from the
oracle.  Nothing in the shipped package imports this file.
"""
import numpy as np
import torch
from scipy.linalg import pinv

from oracle import sktensor_oracle as orc


class NumpyBackend(object):
    def to_device(self, a):
        return torch.from_numpy(np.array(a, dtype=np.float64, order='C', copy=True))

    def to_host(self, x):
        return x.detach().numpy().copy()

    def empty(self, shape, dtype=np.float64):
        return torch.zeros(tuple(int(s) for s in shape), dtype=getattr(torch, np.dtype(dtype).name))

    zeros = empty

    def upload_dense(self, X):
        return np.array(X, dtype=np.float64, order='C', copy=True)

    def upload_coo(self, subs, vals, shape):
        return (tuple(np.asarray(s, dtype=np.int64) for s in subs), np.asarray(vals, dtype=np.float64), tuple(shape))

    @staticmethod
    def _np(U):
        return [None if u is None else u.numpy() for u in U]

    def mttkrp_dense(self, Xd, shape, U, R, n, out):
        Un = self._np(U)
        out.copy_(torch.from_numpy(orc.mttkrp_dense(np.asarray(Xd).reshape(tuple(int(d) for d in shape)), Un, n)))
        return out

    def kr_reduce(self, T, dims, W, R, n, out):
        # the partial tensor with its trailing rank axis, one rank column at a time
        t = T.numpy().reshape(tuple(int(d) for d in dims) + (R,))
        Wn = self._np(W)
        res = np.empty((int(dims[n]), R))
        for r in range(R):
            v = t[..., r]
            for m in range(len(dims) - 1, -1, -1):
                if m != n:
                    v = np.tensordot(v, Wn[m][:, r], axes=([m], [0]))
            res[:, r] = v
        out.copy_(torch.from_numpy(res))
        return out

    def mttkrp_coo(self, coo, U, R, n, out):
        subs, vals, shape = coo
        out.copy_(torch.from_numpy(orc.mttkrp_coo(subs, vals, shape, self._np(U), n)))
        return out

    def sumsq(self, x):
        return torch.tensor([float((np.asarray(x) ** 2).sum())], dtype=torch.float64)

    def coo_sumsq(self, coo):
        return torch.tensor([float((coo[1] ** 2).sum())], dtype=torch.float64)

    def gram(self, U, out):
        u = U.numpy()
        out.copy_(torch.from_numpy(u.T.dot(u)))
        return out

    def hadamard_pinv(self, G, n, P, info):
        g = G.numpy()
        Y = np.ones(g.shape[1:])
        for m in range(g.shape[0]):
            if m != n:
                Y = Y * g[m]
        if not np.isfinite(Y).all():
            info[0] = -1
            P.fill_(float('nan'))
        else:
            P.copy_(torch.from_numpy(pinv(Y)))
            info[0] = int(np.linalg.matrix_rank(Y))
        return P, info

    def solve_stats(self, M, P, first, Unew, colstat):
        u = M.numpy().dot(P.numpy())
        Unew.copy_(torch.from_numpy(u))
        if u.shape[0] == 0:
            stat = np.zeros(u.shape[1]) if first else np.full(u.shape[1], -np.finfo(float).max)
        else:
            stat = (u ** 2).sum(axis=0) if first else u.max(axis=0)
        colstat.copy_(torch.from_numpy(stat))
        return Unew, colstat

    def normalise_gram(self, U, colstat, first, lam, G, Mip, ip):
        c = colstat.numpy()
        l = np.sqrt(c) if first else np.where(c < 1, 1.0, c)
        lam.copy_(torch.from_numpy(l))
        u = U.numpy() / l
        U.copy_(torch.from_numpy(u))
        G.copy_(torch.from_numpy(u.T.dot(u)))
        if Mip is not None:
            ip[0] = float((l * u * Mip.numpy()).sum())
        return lam, G, ip

    def cp_fit(self, G, lam, ip, normX2, out):
        l = lam.numpy()
        coef = np.outer(l, l)
        for g in G.numpy():
            coef = coef * g
        normP2 = coef.sum()
        nx2 = float(normX2[0])
        out[0] = 1.0 - (nx2 + normP2 - 2.0 * float(ip[0])) / nx2
        out[1] = normP2
        return out

    # the fused small-factor update (cpals_fused.cu): same arithmetic, one call
    def update_fused_fits(self, rows, R):
        return int(rows) * int(R) <= 20480 and R <= 64

    def update_fused(self, M, G, n, first, U, lam, info, ip, normX2, fit):
        P = torch.zeros(G.shape[1:], dtype=torch.float64)
        self.hadamard_pinv(G, n, P, info)
        colstat = torch.zeros(G.shape[1], dtype=torch.float64)
        self.solve_stats(M, P, first, U, colstat)
        self.normalise_gram(U, colstat, first, lam, G[n], M if ip is not None else None, ip)
        if fit is not None:
            self.cp_fit(G, lam, ip, normX2, fit)

    # the cluster update with the side-stream pinv (cpals_fused.cu): same arithmetic, synchronous here
    def update_cluster_fits(self, rows, R):
        return int(rows) * int(R) <= 100000 and R <= 64

    def pinv_async(self, G, n, P, info):
        self.hadamard_pinv(G, n, P, info)
        return ('pinv', n)

    def wait_pinv(self, handle):
        assert handle[0] == 'pinv'

    def update_cluster(self, M, P, G, n, first, U, lam, ip, normX2, fit):
        colstat = torch.zeros(G.shape[1], dtype=torch.float64)
        self.solve_stats(M, P, first, U, colstat)
        self.normalise_gram(U, colstat, first, lam, G[n], M if ip is not None else None, ip)
        if fit is not None:
            self.cp_fit(G, lam, ip, normX2, fit)

    # exchange-form update (cpals_fused.cu): record [colstat | Unew^T Unew | <Unew, M>], finish in rank order
    def exchange_record(self, R):
        return R + R * R + 1

    def update_exchange(self, M, P, first, Unew, ex):
        u = M.numpy().dot(P.numpy())
        Unew.copy_(torch.from_numpy(u))
        R = P.shape[0]
        if u.shape[0] == 0:
            stat = np.zeros(R) if first else np.full(R, -np.finfo(float).max)
        else:
            stat = (u ** 2).sum(axis=0) if first else u.max(axis=0)
        rec = np.concatenate([stat, u.T.dot(u).ravel(), [float((u * M.numpy()).sum())]])
        ex.copy_(torch.from_numpy(rec))

    def update_finish(self, ex_all, world, first, U, lam, G, n, ip, normX2, fit):
        R = G.shape[1]
        e = ex_all.numpy()[:world]
        stat = e[:, :R].sum(axis=0) if first else e[:, :R].max(axis=0)
        l = np.sqrt(stat) if first else np.where(stat < 1, 1.0, stat)
        lam.copy_(torch.from_numpy(l))
        if U.shape[0]:
            U.copy_(torch.from_numpy(U.numpy() / l))
        g = e[:, R:R + R * R].sum(axis=0).reshape(R, R) / np.outer(l, l)
        G[n].copy_(torch.from_numpy(g))
        tot = float(e[:, R + R * R].sum())
        if ip is not None:
            ip[0] = tot
        if fit is not None:
            ipt = torch.tensor([tot], dtype=torch.float64)
            self.cp_fit(G, lam, ipt, normX2, fit)

// Fused factor update for small factors: everything cp.als does between two MTTKRPs
// (sktensor/cp.py:144-159) in ONE single-CTA kernel when the factor fits in shared memory:
//
//   Y = hadamard_{m != n} G_m ;  P = pinv(Y)            cp.py:144-147   (pinv.cuh)
//   Unew = M P                                          cp.py:147       (FP64 tensor-core MMA, operands in smem)
//   lambda = column 2-norms (itr 0) | max(col max, 1)   cp.py:149-153
//   U_n = Unew / lambda ;  G_n = U_n^T U_n              cp.py:154, 146  (DMMA again)
//   <P,X> = sum_r lambda_r sum_i U[i,r] M[i,r]          ktensor.py:128-150 via the last-mode MTTKRP
//   fit = 1 - (|X|^2 + |P|^2 - 2<P,X>) / |X|^2          cp.py:157-159, ktensor.py:113-126
//
// For the dense configurations the factors are tiny (512 x 32, 160 x 64): three launches with grid-wide
// "last block" reductions cost ~55 us per mode against ~300 us for the MTTKRP itself; this kernel does the
// same arithmetic in ~10 us.  Tall factors (sparse tensors: 10^5..10^7 rows) keep the streaming kernels of
// cpals.cu.  All reductions are ordered: results are bitwise reproducible.
#include <algorithm>
#include <cfloat>
#include <cstdlib>

#include "common.cuh"
#include "pinv.cuh"

namespace skt {
namespace {

constexpr int FT = 1024;            // threads
constexpr int FW = FT / 32;         // warps
constexpr int QMAX = 10;            // 8x8 output blocks per warp in the solve (covers rows8*R8/64 <= 320)
constexpr int MAXELEMS = 20480;     // rows8 * R8 limit (160 KB of factor in shared memory)

struct FusedParams {
    const double *M;      // rows x R MTTKRP
    double *G;            // [ndim][R][R]; G[mode] is rewritten
    double *U;            // rows x R out
    double *lambda;       // R out
    int *info;            // rank kept by pinv, -1 on non-finite Y
    double *ip;           // <P,X> out (may be NULL)
    const double *normX2; // with fit: |X|^2
    double *fit;          // out[0] = fit, out[1] = |P|^2 (may be NULL)
    long long rows;
    int R, ndim, mode, first_iter;
};

__host__ __device__ inline int round8(int x) { return (x + 7) & ~7; }

struct FusedLayout {
    int R8, LD, rows8;
    size_t ms, ps, pv, cs, gp, lam, red, total;   // offsets in doubles
};

__host__ __device__ inline FusedLayout fused_layout(long long rows, int R) {
    FusedLayout L;
    L.R8 = round8(R);
    L.LD = L.R8 + 4;                       // == 4 or 12 mod 16: conflict-free DMMA fragment loads
    L.rows8 = round8((int)rows);
    size_t o = 0;
    L.ms = o; o += (size_t)L.rows8 * L.LD;          // M, then Unew, then U
    L.ps = o; o += (size_t)L.R8 * L.LD;             // P (zero padded)
    L.pv = o; o += pinv_smem_doubles(R);            // pinv scratch; later: column partials / Gram partials
    L.cs = L.pv;                                    // [nchunk][RP] column-statistic partials (aliases pinv scratch)
    L.gp = L.pv;                                    // Gram partials [S][R8][R8+1]   (aliases pinv scratch)
    const size_t gp_need = (size_t)FW * 64 + 64;    // at most 32 jobs x 64 entries in flight per round
    const size_t cs_need = (size_t)FT;
    const size_t alias = gp_need > cs_need ? gp_need : cs_need;
    if (pinv_smem_doubles(R) < alias) o = L.pv + alias;
    L.lam = o; o += 64;
    L.red = o; o += 64;
    L.total = o;
    return L;
}

__global__ void __launch_bounds__(FT, 1) cp_update_fused_kernel(const __grid_constant__ FusedParams p) {
    extern __shared__ __align__(16) double fsm[];
    const int R = p.R, rows = (int)p.rows;
    const FusedLayout L = fused_layout(p.rows, R);
    const int R8 = L.R8, LD = L.LD, rows8 = L.rows8;
    double *Ms = fsm + L.ms, *Ps = fsm + L.ps, *cs = fsm + L.cs, *gp = fsm + L.gp, *lam = fsm + L.lam, *red = fsm + L.red;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;

    // ---- stage M (zero padded to rows8 x R8) while the pseudo-inverse is being formed
    for (int e = tid; e < rows8 * R8; e += FT) {
        const int i = e / R8, c = e - i * R8;
        Ms[i * LD + c] = (i < rows && c < R) ? __ldg(p.M + (long long)i * R + c) : 0.0;
    }
    for (int e = tid; e < R8 * LD; e += FT) Ps[e] = 0.0;

    // ---- P = pinv(Y) straight into shared memory
    const PinvShared ps = pinv_carve(fsm + L.pv, R);
    const bool bad = pinv_load_y(ps, p.G, p.ndim, R, p.mode);   // contains block barriers: Ms / Ps are visible after it
    if (bad) {
        // the reference raises ValueError('array must not contain infs or NaNs') inside pinv (cp.py:147)
        for (int e = tid; e < rows * R; e += FT) p.U[e] = nan("");
        if (tid == 0) p.info[0] = -1;
        return;
    }
    const int kept = block_pinv(ps, p.G, p.ndim, R, p.mode, [&](int a, int b, double v) { Ps[a * LD + b] = v; });
    if (tid == 0) p.info[0] = kept;
    __syncthreads();

    // ---- Unew = M P : 8x8 output blocks, K = R8, FP64 tensor cores; warp w owns blocks w, w+32, ...
    const int ncb = R8 >> 3, nrb = rows8 >> 3, nblk = nrb * ncb;
    double acc[QMAX][2];
    double ipp = 0.0;
#pragma unroll
    for (int q = 0; q < QMAX; ++q) {
        acc[q][0] = acc[q][1] = 0.0;
        const int blk = warp + FW * q;
        if (blk < nblk) {
            const int ib = blk / ncb, jb = blk - ib * ncb;
            const double *arow = Ms + (ib * 8 + g) * LD + t4;
            const double *bcol = Ps + t4 * LD + jb * 8 + g;
            for (int k0 = 0; k0 < R8; k0 += 4) dmma884(acc[q][0], acc[q][1], arow[k0], bcol[k0 * LD]);
            if (p.ip) {   // <P,X> = sum Unew[i,c] M[i,c]  (lambda_c * U[i,c] == Unew[i,c])
                const double *mrow = Ms + (ib * 8 + g) * LD + jb * 8 + 2 * t4;
                ipp += acc[q][0] * mrow[0] + acc[q][1] * mrow[1];
            }
        }
    }
    __syncthreads();            // every warp is done reading M
#pragma unroll
    for (int q = 0; q < QMAX; ++q) {
        const int blk = warp + FW * q;
        if (blk < nblk) {
            const int ib = blk / ncb, jb = blk - ib * ncb;
            double *urow = Ms + (ib * 8 + g) * LD + jb * 8 + 2 * t4;
            urow[0] = acc[q][0];
            urow[1] = acc[q][1];
        }
    }
    // <P,X>: ordered block sum
    if (p.ip) {
        ipp = warp_sum(ipp);
        if (lane == 0) red[warp] = ipp;
    }
    __syncthreads();
    if (p.ip && tid == 0) {
        double sum = 0.0;
        for (int w = 0; w < FW; ++w) sum += red[w];
        p.ip[0] = sum;
        red[FW] = sum;
    }

    // ---- column statistics: sum of squares (first iteration) or signed maximum, in a fixed order
    {
        const int RP = (R <= 16) ? 16 : (R <= 32 ? 32 : 64);
        const int nchunk = FT / RP;
        const int c = tid % RP, ch = tid / RP;
        const int per = (rows + nchunk - 1) / nchunk;
        double st = p.first_iter ? 0.0 : -DBL_MAX;
        if (c < R) {
            const int i1 = min(rows, (ch + 1) * per);
            for (int i = ch * per; i < i1; ++i) {
                const double v = Ms[i * LD + c];
                st = p.first_iter ? st + v * v : nanmax(st, v);
            }
        }
        cs[ch * RP + c] = st;
        __syncthreads();
        if (tid < R) {
            double sacc = p.first_iter ? 0.0 : -DBL_MAX;
            for (int k = 0; k < nchunk; ++k) {
                const double v = cs[k * RP + tid];
                sacc = p.first_iter ? sacc + v : nanmax(sacc, v);
            }
            const double l = p.first_iter ? sqrt(sacc) : (sacc < 1.0 ? 1.0 : sacc);   // cp.py:150 / cp.py:152-153
            lam[tid] = l;
            p.lambda[tid] = l;
        }
        __syncthreads();
    }

    // ---- U = Unew / lambda (shared + global)
    for (int e = tid; e < rows * R; e += FT) {
        const int i = e / R, c = e - i * R;
        const double v = Ms[i * LD + c] / lam[c];
        Ms[i * LD + c] = v;
        p.U[e] = v;
    }
    __syncthreads();

    // ---- G_n = U^T U : 8x8 blocks x S row parts on the tensor cores, parts summed in order
    {
        const int nb2 = ncb * ncb;
        const int S = max(1, FW / nb2);
        const int rp = (((rows8 + S - 1) / S) + 3) & ~3;   // rows per part, a multiple of the DMMA k step
        const int njobs = nb2 * S;
        double *Gn = p.G + (size_t)p.mode * R * R;
        for (int j0 = 0; j0 < njobs; j0 += FW) {
            const int job = j0 + warp;
            double c0 = 0.0, c1 = 0.0;
            int ab = 0, bb = 0, part = 0;
            if (job < njobs) {
                const int b2 = job % nb2;
                part = job / nb2;
                ab = b2 / ncb; bb = b2 - ab * ncb;
                const int k1 = min(rows8, (part + 1) * rp);
                const double *ap = Ms + t4 * LD + ab * 8 + g;
                const double *bp = Ms + t4 * LD + bb * 8 + g;
                for (int k = part * rp; k < k1; k += 4) dmma884(c0, c1, ap[k * LD], bp[k * LD]);
                double *dst = gp + (size_t)warp * 64 + g * 8 + 2 * t4;
                dst[0] = c0;
                dst[1] = c1;
            }
            __syncthreads();
            // the S parts of one block sit in warps b2, b2 + nb2, ... of this round
            if (S > 1) {
                for (int e = tid; e < nb2 * 64; e += FT) {
                    const int b2 = e >> 6, w = e & 63;
                    double sum = 0.0;
                    for (int s_ = 0; s_ < S; ++s_) sum += gp[(size_t)(s_ * nb2 + b2) * 64 + w];
                    const int a_ = (b2 / ncb) * 8 + (w >> 3), b_ = (b2 % ncb) * 8 + (w & 7);
                    if (a_ < R && b_ < R) Gn[a_ * R + b_] = sum;
                }
            } else {
                for (int e = tid; e < FW * 64; e += FT) {
                    const int jb_ = j0 + (e >> 6), w = e & 63;
                    if (jb_ < njobs) {
                        const int a_ = (jb_ / ncb) * 8 + (w >> 3), b_ = (jb_ % ncb) * 8 + (w & 7);
                        if (a_ < R && b_ < R) Gn[a_ * R + b_] = gp[e];
                    }
                }
            }
            __syncthreads();
        }
    }

    // ---- fit (cp.py:157-159); G_n was written by this block, so it is visible to it after the barrier
    if (p.fit) {
        double sacc = 0.0;
        for (int e = tid; e < R * R; e += FT) {
            const int a = e / R, b = e - a * R;
            double y = lam[a] * lam[b];
            for (int k = 0; k < p.ndim; ++k) y *= p.G[(size_t)k * R * R + e];
            sacc += y;
        }
        sacc = warp_sum(sacc);
        __syncthreads();
        if (lane == 0) red[warp] = sacc;
        __syncthreads();
        if (tid == 0) {
            double normP2 = 0.0;
            for (int w = 0; w < FW; ++w) normP2 += red[w];
            const double nx2 = p.normX2[0];
            p.fit[0] = 1.0 - (nx2 + normP2 - 2.0 * red[FW]) / nx2;
            p.fit[1] = normP2;
        }
    }
}


// =============================================================================================
// Cluster version of the update (everything but the pseudo-inverse): 8 CTAs of one thread-block cluster
// split the factor rows, exchange the R column statistics and the R x R partial Grams through distributed
// shared memory (ordered by CTA rank: deterministic) and finish in a few microseconds.  P = pinv(Y) comes from
// skt_hadamard_pinv, which the engine runs on a side stream WHILE the MTTKRP of the same mode is still
// streaming the tensor (Y only needs the Grams of the other modes), so it leaves the critical path.
// =============================================================================================
constexpr int CL = 8;               // CTAs per cluster (portable maximum)
constexpr int CTH = 512;            // threads per CTA
constexpr int CWARPS = CTH / 32;
constexpr int CMAXSLICE = 12288;    // per-CTA slice limit: rows_per_cta * LD doubles (96 KB)

struct ClusterParams {
    const double *M, *P;
    double *G, *U, *lambda, *ip;
    const double *normX2;
    double *fit;
    long long rows;
    int R, ndim, mode, first_iter;
    double *ex;   // non-NULL: phase-A mode -- write Unew (not normalised) to U and [colstat | Unew^T Unew | <Unew,M>] to ex
    // nparts > 1: M is the sum, in rank order, of the ranks' partial MTTKRPs, read straight from the peers' HBM over
    // NVLink (symmetric memory: Mparts[w] is rank w's buffer mapped into this process) -- the all-reduce of the partial
    // MTTKRP is fused into the staging loop of the update
    const double *const *Mparts;
    int nparts;
    // nslabs >= 1 (and Mparts == NULL): M is the sum of local partial slabs M + s * slab_stride (skt_kr_reduce_slabs)
    int nslabs;
    long long slab_stride;
};

struct ClusterLayout {
    int R8, LD, per;                // per: rows per CTA (multiple of 8)
    size_t ms, us, ps, gp, cst, lam, red, total;
};

__host__ __device__ inline ClusterLayout cluster_layout(long long rows, int R) {
    ClusterLayout L;
    L.R8 = round8(R);
    L.LD = L.R8 + 4;
    const int rows8 = round8((int)rows);
    L.per = ((rows8 / 8 + CL - 1) / CL) * 8;
    if (L.per < 8) L.per = 8;
    size_t o = 0;
    L.ms = o; o += (size_t)L.per * L.LD;
    L.us = o; o += (size_t)L.per * L.LD;
    L.ps = o; o += (size_t)L.R8 * L.LD;          // P, later the summed Gram on rank 0
    L.gp = o; o += ((size_t)L.R8 * L.R8 > (size_t)CTH) ? (size_t)L.R8 * L.R8 : (size_t)CTH;   // partial Gram; first the column-statistic scratch
    L.cst = o; o += 72;                          // [0..63] column statistic partials, [64] <P,X> partial
    L.lam = o; o += 64;
    L.red = o; o += 64;
    L.total = o;
    return L;
}

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(CTH, 1)
    cp_update_cluster_kernel(const __grid_constant__ ClusterParams p) {
    extern __shared__ __align__(16) double csm2[];
    const int R = p.R, rows = (int)p.rows;
    const ClusterLayout L = cluster_layout(p.rows, R);
    const int R8 = L.R8, LD = L.LD, per = L.per;
    double *Ms = csm2 + L.ms, *Us = csm2 + L.us, *Ps = csm2 + L.ps, *gp = csm2 + L.gp, *cst = csm2 + L.cst;
    double *lam = csm2 + L.lam, *red = csm2 + L.red;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    unsigned crank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(crank));
    const int r0 = (int)crank * per;                       // first row of this CTA's slice
    const int nr = max(0, min(per, rows - r0));            // valid rows in it

    auto cluster_sync = [&]() {
        asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    };
    // generic address of `ptr` (own shared memory) as seen in the shared memory of CTA `rank` of the cluster
    auto remote = [&](const double *ptr, unsigned rank) -> const double * {
        uint64_t in = (uint64_t)(uintptr_t)ptr, out;
        asm volatile("mapa.u64 %0, %1, %2;\n" : "=l"(out) : "l"(in), "r"(rank));
        return (const double *)(uintptr_t)out;
    };

    // ---- stage the slice of M and P (zero padded)
    if (p.Mparts) {
        for (int e = tid; e < per * R8; e += CTH) {
            const int i = e / R8, c = e - i * R8;
            double v = 0.0;
            if (i < nr && c < R) {
                const long long off = (long long)(r0 + i) * R + c;
                for (int w = 0; w < p.nparts; ++w) v += p.Mparts[w][off];
            }
            Ms[i * LD + c] = v;
        }
    } else if (p.nslabs > 1) {
        for (int e = tid; e < per * R8; e += CTH) {
            const int i = e / R8, c = e - i * R8;
            double v = 0.0;
            if (i < nr && c < R) {
                const double *src = p.M + (long long)(r0 + i) * R + c;
                for (int sl = 0; sl < p.nslabs; ++sl) v += __ldg(src + sl * p.slab_stride);
            }
            Ms[i * LD + c] = v;
        }
    } else
    for (int e = tid; e < per * R8; e += CTH) {
        const int i = e / R8, c = e - i * R8;
        Ms[i * LD + c] = (i < nr && c < R) ? __ldg(p.M + (long long)(r0 + i) * R + c) : 0.0;
    }
    for (int e = tid; e < R8 * LD; e += CTH) {
        const int a = e / LD, b = e - a * LD;
        Ps[e] = (a < R && b < R) ? __ldg(p.P + a * R + b) : 0.0;
    }
    __syncthreads();

    // ---- Unew = M P on the slice (FP64 tensor cores); <P,X> partial = sum Unew .* M
    const int ncb = R8 >> 3, nblk = (per >> 3) * ncb;
    double ipp = 0.0;
    for (int blk = warp; blk < nblk; blk += CWARPS) {
        const int ib = blk / ncb, jb = blk - ib * ncb;
        const double *arow = Ms + (ib * 8 + g) * LD + t4;
        const double *bcol = Ps + t4 * LD + jb * 8 + g;
        double c0 = 0.0, c1 = 0.0;
        for (int k0 = 0; k0 < R8; k0 += 4) dmma884(c0, c1, arow[k0], bcol[k0 * LD]);
        const int off = (ib * 8 + g) * LD + jb * 8 + 2 * t4;
        Us[off] = c0;
        Us[off + 1] = c1;
        ipp += c0 * Ms[off] + c1 * Ms[off + 1];
    }
    ipp = warp_sum(ipp);
    if (lane == 0) red[warp] = ipp;
    __syncthreads();

    // ---- per-CTA column statistics (fixed order), then the cluster-wide combination in rank order
    {
        const int RP = (R <= 16) ? 16 : (R <= 32 ? 32 : 64);
        const int nchunk = CTH / RP;
        const int c = tid % RP, ch = tid / RP;
        const int perc = (nr + nchunk - 1) / nchunk;
        double st = p.first_iter ? 0.0 : -DBL_MAX;
        if (c < R) {
            const int i1 = min(nr, (ch + 1) * perc);
            for (int i = ch * perc; i < i1; ++i) {
                const double v = Us[i * LD + c];
                st = p.first_iter ? st + v * v : nanmax(st, v);
            }
        }
        double *scr = gp;                                   // [nchunk][RP] scratch (the Gram comes later)
        scr[ch * RP + c] = st;
        __syncthreads();
        if (tid < R) {
            double sacc = p.first_iter ? 0.0 : -DBL_MAX;
            for (int k = 0; k < nchunk; ++k) {
                const double v = scr[k * RP + tid];
                sacc = p.first_iter ? sacc + v : nanmax(sacc, v);
            }
            cst[tid] = sacc;
        }
        if (tid == 0) {
            double sum = 0.0;
            for (int w = 0; w < CWARPS; ++w) sum += red[w];
            cst[64] = sum;
        }
    }
    cluster_sync();
    if (tid < R) {
        double sacc = p.first_iter ? 0.0 : -DBL_MAX;
        for (unsigned k = 0; k < CL; ++k) {
            const double v = remote(cst, k)[tid];
            sacc = p.first_iter ? sacc + v : nanmax(sacc, v);
        }
        double l = p.first_iter ? sqrt(sacc) : (sacc < 1.0 ? 1.0 : sacc);   // cp.py:150 / cp.py:152-153
        if (p.ex) {                      // phase A of a sharded update: the statistic still has to meet the other ranks'
            if (crank == 0) p.ex[tid] = sacc;
            l = 1.0;
        } else if (crank == 0) {
            p.lambda[tid] = l;
        }
        lam[tid] = l;
    }
    __syncthreads();

    // ---- U = Unew / lambda on the slice (shared + global)
    for (int e = tid; e < nr * R; e += CTH) {
        const int i = e / R, c = e - i * R;
        const double v = Us[i * LD + c] / lam[c];
        Us[i * LD + c] = v;
        p.U[(long long)(r0 + i) * R + c] = v;
    }
    __syncthreads();

    // ---- partial Gram of the slice (rows beyond nr are zero)
    {
        const int nb2 = ncb * ncb;
        for (int job = warp; job < nb2; job += CWARPS) {
            const int ab = job / ncb, bb = job - ab * ncb;
            const double *ap = Us + t4 * LD + ab * 8 + g;
            const double *bp = Us + t4 * LD + bb * 8 + g;
            double c0 = 0.0, c1 = 0.0;
            for (int k = 0; k < per; k += 4) dmma884(c0, c1, ap[k * LD], bp[k * LD]);
            double *dst = gp + (ab * 8 + g) * R8 + bb * 8 + 2 * t4;
            dst[0] = c0;
            dst[1] = c1;
        }
    }
    cluster_sync();

    // ---- rank 0: G_n = sum of the partial Grams in rank order, <P,X>, fit
    if (crank == 0) {
        double *Gn = p.ex ? (p.ex + R) : (p.G + (size_t)p.mode * R * R);
        double *Gs = Ps;                                    // P is no longer needed
        for (int e = tid; e < R * R; e += CTH) {
            const int a = e / R, b = e - a * R;
            double sum = 0.0;
            for (unsigned k = 0; k < CL; ++k) sum += remote(gp, k)[a * R8 + b];
            Gn[e] = sum;
            Gs[e] = sum;
        }
        if (tid == 0 && (p.ip || p.ex)) {
            double sum = 0.0;
            for (unsigned k = 0; k < CL; ++k) sum += remote(cst, k)[64];
            if (p.ex) p.ex[R + R * R] = sum;
            else p.ip[0] = sum;
            red[CWARPS] = sum;
        }
        __syncthreads();
        if (p.fit && !p.ex) {
            double sacc = 0.0;
            for (int e = tid; e < R * R; e += CTH) {
                const int a = e / R, b = e - a * R;
                double y = lam[a] * lam[b];
                for (int k = 0; k < p.ndim; ++k) y *= (k == p.mode) ? Gs[e] : __ldg(p.G + (size_t)k * R * R + e);
                sacc += y;
            }
            sacc = warp_sum(sacc);
            if (lane == 0) red[warp] = sacc;
            __syncthreads();
            if (tid == 0) {
                double normP2 = 0.0;
                for (int w = 0; w < CWARPS; ++w) normP2 += red[w];
                const double nx2 = p.normX2[0];
                p.fit[0] = 1.0 - (nx2 + normP2 - 2.0 * red[CWARPS]) / nx2;
                p.fit[1] = normP2;
            }
        }
    }
    cluster_sync();     // the peers' shared memory must outlive rank 0's reads
}


// =============================================================================================
// Exchange form of the update, for a factor whose rows are spread over ranks (the sharded mode) or too tall for
// shared memory (sparse tensors).  Phase A (cluster kernel above with `ex`, or solve_exchange_kernel for tall
// factors) writes Unew = M P and ONE small record  ex = [colstat (R) | Unew^T Unew (R x R) | <Unew, M>];  the
// ranks all-gather their records (one collective instead of three all-reduces);  cp_finish_kernel then forms,
// in rank order,  lambda,  G_n = (sum_w Unew_w^T Unew_w) ./ (lambda lambda^T),  <P,X>,  the fit, and divides the
// local rows by lambda -- a pass that is skipped altogether when every lambda is 1 (the usual case after the
// first sweep, cp.py:152-153).
// =============================================================================================
constexpr int SXT = 256, SXROWS = 64;

template <int TT>
__global__ void __launch_bounds__(SXT) solve_exchange_kernel(const double *__restrict__ M, const double *__restrict__ P,
                                                             long long rows, int R, int first_iter, double *Unew,
                                                             double *ex, double *partials, unsigned int *counter) {
    // 16 x 16 threads; thread (ty, tx) owns Unew[ty + 16a][tx + 16b] of a 64-row tile (a < 4) and, for the Gram,
    // G[ty + 16a'][tx + 16b] (a', b < TT)
    extern __shared__ double sxm[];
    constexpr int PL = 16 * TT + 1;
    double *Ps = sxm;                                  // [R][PL]
    double *Ms = Ps + SKT_MAX_RANK * PL;               // [SXROWS][R + 1]
    double *Us = Ms + SXROWS * (SKT_MAX_RANK + 1);     // [SXROWS][PL]
    double *cs = Us + SXROWS * PL;                     // [16][16 * TT]
    __shared__ double red[SXT / 32];
    const int ML = R + 1;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int nblk = gridDim.x;
    for (int e = tid; e < R * 16 * TT; e += SXT) {
        const int k = e / (16 * TT), c = e % (16 * TT);
        Ps[k * PL + c] = (c < R) ? P[k * R + c] : 0.0;
    }
    double stat[TT], gacc[TT][TT], ip = 0.0;
#pragma unroll
    for (int b = 0; b < TT; ++b) {
        stat[b] = first_iter ? 0.0 : -DBL_MAX;
#pragma unroll
        for (int a = 0; a < TT; ++a) gacc[a][b] = 0.0;
    }
    __syncthreads();
    const long long ntile = (rows + SXROWS - 1) / SXROWS;
    for (long long tile = blockIdx.x; tile < ntile; tile += nblk) {
        const long long r0 = tile * SXROWS;
        const int nr = (int)min((long long)SXROWS, rows - r0);
        for (int e = tid; e < SXROWS * R; e += SXT) {
            const int i = e / R, c = e % R;
            Ms[i * ML + c] = (i < nr) ? M[(r0 + i) * R + c] : 0.0;
        }
        __syncthreads();
        double acc[4][TT];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < TT; ++b) acc[a][b] = 0.0;
        for (int k = 0; k < R; ++k) {
            double m[4], pv[TT];
#pragma unroll
            for (int a = 0; a < 4; ++a) m[a] = Ms[(ty + 16 * a) * ML + k];
#pragma unroll
            for (int b = 0; b < TT; ++b) pv[b] = Ps[k * PL + tx + 16 * b];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < TT; ++b) acc[a][b] += m[a] * pv[b];
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int i = ty + 16 * a;
#pragma unroll
            for (int b = 0; b < TT; ++b) {
                const int c = tx + 16 * b;
                const bool ok = (i < nr) && (c < R);
                const double v = ok ? acc[a][b] : 0.0;
                Us[i * PL + c] = v;
                if (ok) {
                    Unew[(r0 + i) * R + c] = v;
                    ip += v * Ms[i * ML + c];
                    if (first_iter) stat[b] += v * v;
                    else stat[b] = nanmax(stat[b], v);
                }
            }
        }
        __syncthreads();
        for (int i = 0; i < nr; ++i) {
            double ua[TT], ub[TT];
#pragma unroll
            for (int a = 0; a < TT; ++a) ua[a] = Us[i * PL + ty + 16 * a];
#pragma unroll
            for (int b = 0; b < TT; ++b) ub[b] = Us[i * PL + tx + 16 * b];
#pragma unroll
            for (int a = 0; a < TT; ++a)
#pragma unroll
                for (int b = 0; b < TT; ++b) gacc[a][b] += ua[a] * ub[b];
        }
        __syncthreads();
    }
    // per-CTA record: [colstat | Gram | ip], column statistics combined over the 16 row groups in a fixed order
#pragma unroll
    for (int b = 0; b < TT; ++b) cs[ty * 16 * TT + tx + 16 * b] = stat[b];
    __syncthreads();
    const int REC = R + R * R + 1;
    double *mine = partials + (size_t)blockIdx.x * REC;
    if (tid < R) {
        double sacc = first_iter ? 0.0 : -DBL_MAX;
        for (int y = 0; y < 16; ++y) {
            const double v = cs[y * 16 * TT + tid];
            sacc = first_iter ? sacc + v : nanmax(sacc, v);
        }
        mine[tid] = sacc;
    }
#pragma unroll
    for (int a = 0; a < TT; ++a)
#pragma unroll
        for (int b = 0; b < TT; ++b) {
            const int ra = ty + 16 * a, cb = tx + 16 * b;
            if (ra < R && cb < R) mine[R + ra * R + cb] = gacc[a][b];
        }
    ip = warp_sum(ip);
    if ((tid & 31) == 0) red[tid >> 5] = ip;
    __syncthreads();
    if (tid == 0) {
        double sum = 0.0;
        for (int w = 0; w < SXT / 32; ++w) sum += red[w];
        mine[R + R * R] = sum;
    }
    if (last_block_ticket(counter, nblk)) {
        for (int e = tid; e < REC; e += SXT) {
            const bool is_stat = e < R;
            double sacc = (is_stat && !first_iter) ? -DBL_MAX : 0.0;
            for (int b = 0; b < nblk; ++b) {
                const double v = partials[(size_t)b * REC + e];
                sacc = (is_stat && !first_iter) ? nanmax(sacc, v) : sacc + v;
            }
            ex[e] = sacc;
        }
    }
}

// Phase B.  ex_all: `world` records of R + R*R + 1 doubles (rank order).  Every CTA rebuilds lambda (cheap) and
// divides its share of the local rows when some lambda differs from 1; CTA 0 writes lambda, G_mode, <P,X>, the fit.
// record w: ex_parts ? ex_parts[w] (rank w's record in symmetric memory, read over NVLink) : ex_all + w * REC
__device__ __forceinline__ const double *finish_record(const double *ex_all, const double *const *ex_parts, int w, int REC) {
    return ex_parts ? ex_parts[w] : ex_all + (size_t)w * REC;
}

__global__ void __launch_bounds__(256) cp_finish_kernel(const double *__restrict__ ex_all, const double *const *ex_parts,
                                                        int world, int first_iter,
                                                        long long rows, int R, int ndim, int mode, double *U, double *lambda,
                                                        double *G, double *ip_out, const double *__restrict__ normX2,
                                                        double *fit) {
    __shared__ double lam[SKT_MAX_RANK];
    __shared__ double red[8];
    __shared__ int need_div;
    const int tid = threadIdx.x;
    const int REC = R + R * R + 1;
    if (tid == 0) need_div = 0;
    __syncthreads();
    if (tid < R) {
        double sacc = first_iter ? 0.0 : -DBL_MAX;
        for (int w = 0; w < world; ++w) {
            const double v = finish_record(ex_all, ex_parts, w, REC)[tid];
            sacc = first_iter ? sacc + v : nanmax(sacc, v);
        }
        const double l = first_iter ? sqrt(sacc) : (sacc < 1.0 ? 1.0 : sacc);   // cp.py:150 / cp.py:152-153
        lam[tid] = l;
        if (l != 1.0) need_div = 1;
        if (blockIdx.x == 0) lambda[tid] = l;
    }
    __syncthreads();
    if (need_div) {
        const long long n = rows * R;
        for (long long e = blockIdx.x * 256LL + tid; e < n; e += (long long)gridDim.x * 256) U[e] = U[e] / lam[e % R];
    }
    if (blockIdx.x != 0) return;
    double *Gn = G + (size_t)mode * R * R;
    double sacc = 0.0;
    for (int e = tid; e < R * R; e += 256) {
        const int a = e / R, b = e - a * R;
        double g = 0.0;
        for (int w = 0; w < world; ++w) g += finish_record(ex_all, ex_parts, w, REC)[R + e];
        g = g / (lam[a] * lam[b]);
        Gn[e] = g;
        if (fit) {
            double y = lam[a] * lam[b] * g;
            for (int k = 0; k < ndim; ++k)
                if (k != mode) y *= G[(size_t)k * R * R + e];
            sacc += y;
        }
    }
    double ip = 0.0;
    if (tid == 0 && (ip_out || fit)) {
        for (int w = 0; w < world; ++w) ip += finish_record(ex_all, ex_parts, w, REC)[R + R * R];
        if (ip_out) ip_out[0] = ip;
    }
    if (fit) {
        sacc = warp_sum(sacc);
        if ((tid & 31) == 0) red[tid >> 5] = sacc;
        __syncthreads();
        if (tid == 0) {
            double normP2 = 0.0;
            for (int w = 0; w < 8; ++w) normP2 += red[w];
            const double nx2 = normX2[0];
            fit[0] = 1.0 - (nx2 + normP2 - 2.0 * ip) / nx2;
            fit[1] = normP2;
        }
    }
}

}  // namespace
}  // namespace skt

using namespace skt;

extern "C" int skt_cp_update_fused_fits(int64_t rows, int rank) {
    if (rows < 1 || rank < 1 || rank > SKT_MAX_RANK || rows > (1 << 20)) return 0;
    const FusedLayout L = fused_layout(rows, rank);
    if ((long long)L.rows8 * L.R8 > MAXELEMS) return 0;
    if ((long long)(L.rows8 / 8) * (L.R8 / 8) > (long long)QMAX * FW) return 0;
    return (L.total * sizeof(double) <= smem_optin_bytes()) ? 1 : 0;
}

extern "C" int skt_cp_update_fused(const double *M_d, int64_t rows, int rank, int ndim, int mode, int first_iter,
                                   double *G_d, double *U_d, double *lambda_d, int *info_d, double *ip_d,
                                   const double *normX2_d, double *fit_d, skt_stream stream) {
    SKT_REQUIRE(M_d && G_d && U_d && lambda_d && info_d, SKT_EINVAL, "NULL argument");
    SKT_REQUIRE(ndim >= 1 && ndim <= SKT_MAX_NDIM && mode >= 0 && mode < ndim, SKT_EINVAL, "bad ndim/mode");
    SKT_REQUIRE(fit_d == nullptr || (ip_d != nullptr && normX2_d != nullptr), SKT_EINVAL, "the fit needs ip_d and normX2_d");
    SKT_REQUIRE(skt_cp_update_fused_fits(rows, rank), SKT_EUNSUPPORTED,
                "factor of %lld x %d does not fit the fused update kernel", (long long)rows, rank);
    const FusedLayout L = fused_layout(rows, rank);
    const size_t smem = L.total * sizeof(double);
    static unsigned long long configured = 0;   /* bit d: configured on device d */
    if (skt::first_use_on_device(configured)) {
        SKT_CUDA(cudaFuncSetAttribute(cp_update_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_optin_bytes()));
    }
    FusedParams p;
    p.M = M_d; p.G = G_d; p.U = U_d; p.lambda = lambda_d; p.info = info_d; p.ip = ip_d; p.normX2 = normX2_d; p.fit = fit_d;
    p.rows = rows; p.R = rank; p.ndim = ndim; p.mode = mode; p.first_iter = first_iter;
    cp_update_fused_kernel<<<1, FT, smem, (cudaStream_t)stream>>>(p);
    SKT_LAUNCH_CHECK();
    return SKT_OK;
}

extern "C" int skt_cp_update_cluster_fits(int64_t rows, int rank) {
    if (rows < 1 || rank < 1 || rank > SKT_MAX_RANK || rows > (1 << 20)) return 0;
    const ClusterLayout L = cluster_layout(rows, rank);
    if ((long long)L.per * L.LD > CMAXSLICE) return 0;
    return (L.total * sizeof(double) <= smem_optin_bytes()) ? 1 : 0;
}

static int update_cluster_impl(const double *M_d, int nslabs, const double *const *Mparts_d, int nparts, const double *P_d,
                               int64_t rows, int rank, int ndim, int mode, int first_iter, double *G_d, double *U_d,
                               double *lambda_d, double *ip_d, const double *normX2_d, double *fit_d, skt_stream stream) {
    SKT_REQUIRE((M_d || (Mparts_d && nparts >= 1)) && P_d && G_d && U_d && lambda_d, SKT_EINVAL, "NULL argument");
    SKT_REQUIRE(ndim >= 1 && ndim <= SKT_MAX_NDIM && mode >= 0 && mode < ndim, SKT_EINVAL, "bad ndim/mode");
    SKT_REQUIRE(fit_d == nullptr || (ip_d != nullptr && normX2_d != nullptr), SKT_EINVAL, "the fit needs ip_d and normX2_d");
    SKT_REQUIRE(skt_cp_update_cluster_fits(rows, rank), SKT_EUNSUPPORTED,
                "factor of %lld x %d does not fit the cluster update kernel", (long long)rows, rank);
    const ClusterLayout L = cluster_layout(rows, rank);
    const size_t smem = L.total * sizeof(double);
    static unsigned long long configured = 0;   /* bit d: configured on device d */
    if (skt::first_use_on_device(configured)) {
        SKT_CUDA(cudaFuncSetAttribute(cp_update_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_optin_bytes()));
    }
    ClusterParams p;
    p.M = M_d; p.P = P_d; p.G = G_d; p.U = U_d; p.lambda = lambda_d; p.ip = ip_d; p.normX2 = normX2_d; p.fit = fit_d;
    p.rows = rows; p.R = rank; p.ndim = ndim; p.mode = mode; p.first_iter = first_iter; p.ex = nullptr; p.Mparts = Mparts_d; p.nparts = nparts; p.nslabs = nslabs; p.slab_stride = (long long)rows * rank;
    cp_update_cluster_kernel<<<CL, CTH, smem, (cudaStream_t)stream>>>(p);
    SKT_LAUNCH_CHECK();
    return SKT_OK;
}

extern "C" size_t skt_cp_exchange_record(int rank) { return (size_t)rank + (size_t)rank * rank + 1; }

extern "C" size_t skt_cp_update_exchange_workspace(int64_t rows, int rank) {
    const long long ntile = (rows + SXROWS - 1) / SXROWS;
    const long long nb = std::max<long long>(1, std::min<long long>(ntile, 2LL * sm_count()));
    return 256 + align_up((size_t)nb * skt_cp_exchange_record(rank) * sizeof(double), 256);
}

extern "C" int skt_cp_update_exchange(const double *M_d, const double *P_d, int64_t rows, int rank, int first_iter,
                                      double *Unew_d, double *ex_d, void *ws_d, size_t ws_bytes, skt_stream stream) {
    SKT_REQUIRE(P_d && ex_d && rows >= 0 && ((M_d && Unew_d) || rows == 0), SKT_EINVAL, "NULL argument or negative rows");
    SKT_REQUIRE(rank >= 1 && rank <= SKT_MAX_RANK, SKT_EUNSUPPORTED, "rank %d outside [1, %d]", rank, SKT_MAX_RANK);
    cudaStream_t st = (cudaStream_t)stream;
    if (rows > 0 && skt_cp_update_cluster_fits(rows, rank)) {      // small slab: the cluster kernel in phase-A mode
        const ClusterLayout L = cluster_layout(rows, rank);
        const size_t smem = L.total * sizeof(double);
        static unsigned long long configured = 0;   /* bit d: configured on device d */
        if (skt::first_use_on_device(configured)) {
            SKT_CUDA(cudaFuncSetAttribute(cp_update_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_optin_bytes()));
        }
        ClusterParams p;
        p.M = M_d; p.P = P_d; p.G = nullptr; p.U = Unew_d; p.lambda = nullptr; p.ip = nullptr; p.normX2 = nullptr; p.fit = nullptr;
        p.rows = rows; p.R = rank; p.ndim = 1; p.mode = 0; p.first_iter = first_iter; p.ex = ex_d; p.Mparts = nullptr; p.nparts = 1; p.nslabs = 1; p.slab_stride = 0;
        cp_update_cluster_kernel<<<CL, CTH, smem, st>>>(p);
        SKT_LAUNCH_CHECK();
        return SKT_OK;
    }
    const size_t need = skt_cp_update_exchange_workspace(rows, rank);
    SKT_REQUIRE(ws_d && ws_bytes >= need, SKT_EWORKSPACE, "workspace too small: need %zu bytes, got %zu", need, ws_bytes);
    unsigned int *counter = (unsigned int *)ws_d;
    double *partials = (double *)((char *)ws_d + 256);
    SKT_CUDA(cudaMemsetAsync(counter, 0, 256, st));
    const long long ntile = (rows + SXROWS - 1) / SXROWS;
    const int nb = (int)std::max<long long>(1, std::min<long long>(ntile, 2LL * sm_count()));
    const int tt = (rank + 15) / 16;
    const size_t smem = (size_t)(SKT_MAX_RANK * (16 * tt + 1) + SXROWS * (SKT_MAX_RANK + 1) + SXROWS * (16 * tt + 1) + 16 * 16 * tt) * sizeof(double);
#define SX(TT_)                                                                                                      \
    do {                                                                                                             \
        static unsigned long long cfg = 0;   /* bit d: configured on device d */                                                                                     \
        if (skt::first_use_on_device(cfg)) {                                                                                                  \
            SKT_CUDA(cudaFuncSetAttribute(solve_exchange_kernel<TT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        }                                                                                                            \
        solve_exchange_kernel<TT_><<<nb, SXT, smem, st>>>(M_d, P_d, rows, rank, first_iter, Unew_d, ex_d, partials, counter); \
    } while (0)
    switch (tt) {
        case 1: SX(1); break;
        case 2: SX(2); break;
        case 3: SX(3); break;
        default: SX(4); break;
    }
#undef SX
    SKT_LAUNCH_CHECK();
    return SKT_OK;
}

static int update_finish_impl(const double *ex_all_d, const double *const *ex_parts_d, int world, int first_iter,
                              int64_t rows, int rank, int ndim, int mode, double *U_d, double *lambda_d, double *G_d,
                              double *ip_d, const double *normX2_d, double *fit_d, skt_stream stream) {
    SKT_REQUIRE((ex_all_d || ex_parts_d) && lambda_d && G_d && world >= 1 && rows >= 0 && (U_d || rows == 0), SKT_EINVAL, "NULL argument");
    SKT_REQUIRE(rank >= 1 && rank <= SKT_MAX_RANK, SKT_EUNSUPPORTED, "rank %d outside [1, %d]", rank, SKT_MAX_RANK);
    SKT_REQUIRE(ndim >= 1 && ndim <= SKT_MAX_NDIM && mode >= 0 && mode < ndim, SKT_EINVAL, "bad ndim/mode");
    SKT_REQUIRE(fit_d == nullptr || normX2_d != nullptr, SKT_EINVAL, "the fit needs normX2_d");
    const long long n = (long long)rows * rank;
    const int nb = (int)std::max<long long>(1, std::min<long long>((n + 256 * 8 - 1) / (256 * 8), 4LL * sm_count()));
    cp_finish_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(ex_all_d, ex_parts_d, world, first_iter, rows, rank, ndim, mode, U_d, lambda_d,
                                                         G_d, ip_d, normX2_d, fit_d);
    SKT_LAUNCH_CHECK();
    return SKT_OK;
}

extern "C" int skt_cp_update_cluster(const double *M_d, const double *P_d, int64_t rows, int rank, int ndim, int mode,
                                     int first_iter, double *G_d, double *U_d, double *lambda_d, double *ip_d,
                                     const double *normX2_d, double *fit_d, skt_stream stream) {
    return update_cluster_impl(M_d, 1, nullptr, 1, P_d, rows, rank, ndim, mode, first_iter, G_d, U_d, lambda_d, ip_d, normX2_d,
                               fit_d, stream);
}

extern "C" int skt_cp_update_cluster_parts(const double *const *Mparts_d, int nparts, const double *P_d, int64_t rows,
                                           int rank, int ndim, int mode, int first_iter, double *G_d, double *U_d,
                                           double *lambda_d, double *ip_d, const double *normX2_d, double *fit_d,
                                           skt_stream stream) {
    SKT_REQUIRE(Mparts_d && nparts >= 1 && nparts <= 64, SKT_EINVAL, "bad partial-buffer list");
    return update_cluster_impl(nullptr, 1, Mparts_d, nparts, P_d, rows, rank, ndim, mode, first_iter, G_d, U_d, lambda_d, ip_d,
                               normX2_d, fit_d, stream);
}

extern "C" int skt_cp_update_finish(const double *ex_all_d, int world, int first_iter, int64_t rows, int rank, int ndim,
                                    int mode, double *U_d, double *lambda_d, double *G_d, double *ip_d,
                                    const double *normX2_d, double *fit_d, skt_stream stream) {
    return update_finish_impl(ex_all_d, nullptr, world, first_iter, rows, rank, ndim, mode, U_d, lambda_d, G_d, ip_d, normX2_d,
                              fit_d, stream);
}

extern "C" int skt_cp_update_finish_parts(const double *const *ex_parts_d, int world, int first_iter, int64_t rows, int rank,
                                          int ndim, int mode, double *U_d, double *lambda_d, double *G_d, double *ip_d,
                                          const double *normX2_d, double *fit_d, skt_stream stream) {
    SKT_REQUIRE(ex_parts_d && world >= 1 && world <= 64, SKT_EINVAL, "bad record list");
    return update_finish_impl(nullptr, ex_parts_d, world, first_iter, rows, rank, ndim, mode, U_d, lambda_d, G_d, ip_d,
                              normX2_d, fit_d, stream);
}

extern "C" int skt_cp_update_cluster_slabs(const double *slabs_d, int nslabs, const double *P_d, int64_t rows, int rank,
                                           int ndim, int mode, int first_iter, double *G_d, double *U_d, double *lambda_d,
                                           double *ip_d, const double *normX2_d, double *fit_d, skt_stream stream) {
    SKT_REQUIRE(slabs_d && nslabs >= 1 && nslabs <= 4096, SKT_EINVAL, "bad slab list");
    return update_cluster_impl(slabs_d, nslabs, nullptr, 1, P_d, rows, rank, ndim, mode, first_iter, G_d, U_d, lambda_d, ip_d,
                               normX2_d, fit_d, stream);
}

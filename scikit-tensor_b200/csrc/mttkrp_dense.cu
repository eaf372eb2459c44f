// Dense FP64 MTTKRP for sm_100a -- the kernel behind dtensor.uttkrp
// (reference: sktensor/dtensor.py:163-167 = unfold (dtensor.py:103-150) . khatrirao (core.py:333-378)).
//
// Formulation (no unfolding, no Khatri-Rao matrix).  For output mode n pick the contraction
// mode c = N-1 (c = 0 when n == N-1).  Every other index forms the "row space", viewed as
// (l, i_n, t): l = the modes left of n, t = the modes between n and c.  Then
//
//     acc[row, r]  = sum_kc X[row, kc] * U_c[kc, r]                 (FP64 tensor-core GEMM)
//     M[i_n, r]    = sum_{l,t} acc[(l,i_n,t), r] * prod_{m != n,c} U_m[i_m, r]   (epilogue)
//
// so the flop count stays 2*R*prod(I) and X is streamed exactly once.  A CTA owns a 128-row
// tile (a power-of-two box bl x bi x bt in (l, i_n, t)), streams it in BK=32 slices through a
// multi-stage cp.async ring, multiplies with DMMA.8x8x4, applies the Khatri-Rao weights to
// the accumulators in registers at the end of each tile, and keeps a second register
// accumulator across all tiles of its work item.  Rows of equal i_n are then summed through
// shared memory in a fixed order and written either straight to M or to a per-split partial
// slab that a second tiny kernel adds up -- no atomics, bitwise reproducible.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include <cuda.h>

#include "common.cuh"

namespace skt {

namespace {

constexpr int BM = 128;        // rows per tile
constexpr int BK = 32;         // contraction slice per pipeline stage
constexpr int NTHREADS = 256;  // 8 warps, 16 tile rows each
constexpr int XS_K = BK + 4;   // row stride (doubles) of a K-major X stage  (== 4 mod 16: conflict-free DMMA loads)
constexpr int XS_M = BM + 4;   // row stride of an M-major X stage
constexpr int ZS = 36;         // row stride of the flush staging tile (32 columns + 4)

struct WeightMode {
    const double *U;  // factor of a mode that is neither the output nor the contraction mode
    int dim;          // its length
    int div;          // index = (space_index / div) % dim
    int space;        // 0: decomposes l, 1: decomposes t
};

struct DenseParams {
    const double *X;
    const double *Uc;   // factor of the contraction mode (Kc x ld), pre-offset to the first column of this pass
    double *out;        // M (nsplit == 1) or the partial slabs [nsplit][In][ld]
    long long out_split_stride;
    long long NR;       // L*In*T = number of rows (the kc stride of an M-major X)
    int R;              // columns handled by this pass (<= 8*NB)
    int ld;             // row stride of every factor and of the output
    int Kc, In, L, T;
    int lg_bt, lg_bi, lg_bl;
    int nit, ntt, nred;  // output-row tiles; t tiles; reduction tiles per output-row tile (= nlt*ntt)
    int nsplit, chunk, nitems;
    int nk;
    int vecX, vecU;
    int nw;
    // TMA kernel: stream-K schedule.  CTA b owns the contiguous tile range [b*ntiles/grid, (b+1)*ntiles/grid) of
    // the linearised (output-row tile, reduction tile) space and flushes when the output-row tile changes.
    long long ntiles;      // nit * nred
    double *slab;          // partial outputs: 2 slots of slab_stride doubles per CTA (first / last partial segment)
    long long slab_stride; // bi * ld
    int direct;          // every tile row is its own output row (L == T == 1): the TMA kernel stores accumulators straight to out
    WeightMode w[SKT_MAX_NDIM];
};

template <int NB, bool MMAJOR>
__global__ void __launch_bounds__(NTHREADS, 1) mttkrp_dense_kernel(const __grid_constant__ DenseParams p) {
    constexpr int STAGES = (NB <= 4) ? 4 : 3;
    constexpr int RS = 8 * NB + 4;
    constexpr int XSTAGE = MMAJOR ? BK * XS_M : BM * XS_K;
    constexpr int STAGE = XSTAGE + BK * RS;
    constexpr int NROWOFF = MMAJOR ? 2 : 8;

    extern __shared__ __align__(16) double smem[];
    double *Z = smem + STAGES * STAGE;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int bt = 1 << p.lg_bt, bi = 1 << p.lg_bi, bl = 1 << p.lg_bl;

    // slot q of tile (it, lt, tt) -> (l, i, t) and validity
    auto decompose = [&](int q, int it, int lt, int tt, int &l, int &i, int &t) -> bool {
        const int t_off = q & (bt - 1);
        const int i_off = (q >> p.lg_bt) & (bi - 1);
        const int l_off = q >> (p.lg_bt + p.lg_bi);
        t = tt * bt + t_off;
        i = it * bi + i_off;
        l = lt * bl + l_off;
        return (t < p.T) && (i < p.In) && (l < p.L);
    };
    auto item_range = [&](int item, int &t0, int &t1) {
        const int s = item % p.nsplit;
        t0 = s * p.chunk;
        t1 = min(t0 + p.chunk, p.nred);
    };

    // ------------------------------------------------------------------ loader state
    int l_item = blockIdx.x, l_tile = 0, l_tile_end = 0, l_ks = 0;
    bool l_active = l_item < p.nitems;
    long long rowoff[NROWOFF];  // element offset of each row this thread copies; -1 = padding row

    auto setup_rows = [&]() {
        const int it = l_item / p.nsplit;
        const int lt = l_tile / p.ntt, tt = l_tile % p.ntt;
#pragma unroll
        for (int j = 0; j < NROWOFF; ++j) {
            const int q = MMAJOR ? (2 * (tid & 63) + j) : ((tid >> 4) + 16 * j);
            int l, i, t;
            const bool ok = decompose(q, it, lt, tt, l, i, t);
            const long long rho = ((long long)l * p.In + i) * p.T + t;
            rowoff[j] = ok ? (MMAJOR ? rho : rho * p.Kc) : -1;
        }
    };
    if (l_active) {
        item_range(l_item, l_tile, l_tile_end);
        setup_rows();
    }

    auto loader_issue = [&](int st) {
        if (l_active) {
            double *Xs = smem + st * STAGE;
            double *Us = Xs + XSTAGE;
            const int k0 = l_ks * BK;
            if (!MMAJOR) {
                const int col = 2 * (tid & 15);
                const int kc = k0 + col;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int q = (tid >> 4) + 16 * j;
                    double *dst = Xs + q * XS_K + col;
                    const long long off = rowoff[j];
                    if (p.vecX) {
                        const bool ok = (off >= 0) && (kc < p.Kc);
                        cp_async16(dst, p.X + (ok ? off + kc : 0), ok ? 16 : 0);
                    } else {
                        const bool ok0 = (off >= 0) && (kc < p.Kc);
                        const bool ok1 = (off >= 0) && (kc + 1 < p.Kc);
                        cp_async8(dst, p.X + (ok0 ? off + kc : 0), ok0 ? 8 : 0);
                        cp_async8(dst + 1, p.X + (ok1 ? off + kc + 1 : 0), ok1 ? 8 : 0);
                    }
                }
            } else {
                const int q = 2 * (tid & 63);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int kr = (tid >> 6) + 4 * j;
                    const int kc = k0 + kr;
                    double *dst = Xs + kr * XS_M + q;
                    const long long koff = (long long)kc * p.NR;
                    if (p.vecX) {
                        const bool ok = (rowoff[0] >= 0) && (kc < p.Kc);
                        cp_async16(dst, p.X + (ok ? koff + rowoff[0] : 0), ok ? 16 : 0);
                    } else {
                        const bool ok0 = (rowoff[0] >= 0) && (kc < p.Kc);
                        const bool ok1 = (rowoff[1] >= 0) && (kc < p.Kc);
                        cp_async8(dst, p.X + (ok0 ? koff + rowoff[0] : 0), ok0 ? 8 : 0);
                        cp_async8(dst + 1, p.X + (ok1 ? koff + rowoff[1] : 0), ok1 ? 8 : 0);
                    }
                }
            }
            // factor slice of the contraction mode: BK rows x 8*NB columns (zero beyond R / Kc)
            for (int c = tid; c < BK * 4 * NB; c += NTHREADS) {
                const int kr = c / (4 * NB);
                const int col = 2 * (c % (4 * NB));
                const int kc = k0 + kr;
                double *dst = Us + kr * RS + col;
                const double *src = p.Uc + (long long)kc * p.ld + col;
                if (p.vecU) {
                    const bool ok = (kc < p.Kc) && (col < p.R);
                    cp_async16(dst, ok ? src : p.Uc, ok ? 16 : 0);
                } else {
                    const bool ok0 = (kc < p.Kc) && (col < p.R);
                    const bool ok1 = (kc < p.Kc) && (col + 1 < p.R);
                    cp_async8(dst, ok0 ? src : p.Uc, ok0 ? 8 : 0);
                    cp_async8(dst + 1, ok1 ? src + 1 : p.Uc, ok1 ? 8 : 0);
                }
            }
            // advance to the next (item, tile, k-slice) of this CTA's static schedule
            if (++l_ks == p.nk) {
                l_ks = 0;
                if (++l_tile == l_tile_end) {
                    l_item += gridDim.x;
                    if (l_item < p.nitems) item_range(l_item, l_tile, l_tile_end);
                    else l_active = false;
                }
                if (l_active) setup_rows();
            }
        }
        cp_async_commit();
    };

    // ------------------------------------------------------------------ consumer state
    double acc[2][NB][2], out[2][NB][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
            acc[a][nb][0] = acc[a][nb][1] = 0.0;
            out[a][nb][0] = out[a][nb][1] = 0.0;
        }
    int c_item = blockIdx.x, c_tile = 0, c_tile_end = 0, c_ks = 0;
    if (c_item < p.nitems) item_range(c_item, c_tile, c_tile_end);

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) loader_issue(s);

    int stage = 0;
    while (c_item < p.nitems) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        loader_issue((stage + STAGES - 1) % STAGES);

        // ---- 128 x 8NB x 32 FP64 tensor-core block product on the current stage
        {
            const double *Xs = smem + stage * STAGE;
            const double *Us = Xs + XSTAGE;
#pragma unroll
            for (int kk = 0; kk < BK; kk += 4) {
                double a0, a1;
                if (MMAJOR) {
                    a0 = Xs[(kk + t4) * XS_M + 16 * warp + g];
                    a1 = Xs[(kk + t4) * XS_M + 16 * warp + 8 + g];
                } else {
                    a0 = Xs[(16 * warp + g) * XS_K + kk + t4];
                    a1 = Xs[(16 * warp + 8 + g) * XS_K + kk + t4];
                }
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) {
                    const double b = Us[(kk + t4) * RS + 8 * nb + g];
                    dmma884(acc[0][nb][0], acc[0][nb][1], a0, b);
                    dmma884(acc[1][nb][0], acc[1][nb][1], a1, b);
                }
            }
        }
        stage = (stage + 1) % STAGES;

        if (++c_ks == p.nk) {
            c_ks = 0;
            // ---- tile epilogue: out += acc * prod_m U_m[i_m(row), col]   (the Khatri-Rao rows, on the fly)
            {
                const int it = c_item / p.nsplit;
                const int lt = c_tile / p.ntt, tt = c_tile % p.ntt;
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    const int q = 16 * warp + 8 * a + g;
                    int l, i, t;
                    const bool ok = decompose(q, it, lt, tt, l, i, t);
                    double wgt[NB][2];
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb) wgt[nb][0] = wgt[nb][1] = ok ? 1.0 : 0.0;
                    if (ok) {
                        for (int m = 0; m < p.nw; ++m) {
                            const int sp = p.w[m].space ? t : l;
                            const int idx = (sp / p.w[m].div) % p.w[m].dim;
                            const double *row = p.w[m].U + (long long)idx * p.ld;
#pragma unroll
                            for (int nb = 0; nb < NB; ++nb) {
                                const int col = 8 * nb + 2 * t4;
                                if (col < p.R) wgt[nb][0] *= __ldg(row + col);
                                if (col + 1 < p.R) wgt[nb][1] *= __ldg(row + col + 1);
                            }
                        }
                    }
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb) {
                        out[a][nb][0] += ok ? wgt[nb][0] * acc[a][nb][0] : 0.0;
                        out[a][nb][1] += ok ? wgt[nb][1] * acc[a][nb][1] : 0.0;
                        acc[a][nb][0] = acc[a][nb][1] = 0.0;
                    }
                }
            }
            if (++c_tile == c_tile_end) {
                // ---- item flush: sum rows of equal i_n in a fixed order, write bi x R outputs
                const int it = c_item / p.nsplit, s = c_item % p.nsplit;
                double *dst = p.out + (long long)s * p.out_split_stride;
                constexpr int NH = (NB + 3) / 4;
#pragma unroll
                for (int h = 0; h < NH; ++h) {
#pragma unroll
                    for (int a = 0; a < 2; ++a) {
                        const int q = 16 * warp + 8 * a + g;
#pragma unroll
                        for (int nn = 0; nn < 4; ++nn) {
                            const int nb = 4 * h + nn;
                            if (nb < NB) {
                                Z[q * ZS + 8 * nn + 2 * t4] = out[a][nb][0];
                                Z[q * ZS + 8 * nn + 2 * t4 + 1] = out[a][nb][1];
                                out[a][nb][0] = out[a][nb][1] = 0.0;
                            }
                        }
                    }
                    __syncthreads();
                    for (int o = tid; o < bi * 32; o += NTHREADS) {
                        const int i_off = o >> 5, c = o & 31;
                        const int col = 32 * h + c;
                        const int i = it * bi + i_off;
                        if (col < p.R && col < 8 * NB && i < p.In) {
                            double sum = 0.0;
                            for (int l_off = 0; l_off < bl; ++l_off)
                                for (int t_off = 0; t_off < bt; ++t_off) {
                                    const int q = (((l_off << p.lg_bi) + i_off) << p.lg_bt) + t_off;
                                    sum += Z[q * ZS + c];
                                }
                            dst[(long long)i * p.ld + col] = sum;
                        }
                    }
                    __syncthreads();
                }
                c_item += gridDim.x;
                if (c_item < p.nitems) item_range(c_item, c_tile, c_tile_end);
            }
        }
    }
    cp_async_wait<0>();
}

// =============================================================================================
// TMA variant (the production path whenever the tensor's rows are 16-byte aligned).
//
// X tiles are fetched by the TMA engine (cp.async.bulk.tensor, SASS UTMALDG) from a tiled
// tensor map with 128-byte swizzle; one elected thread issues the bulk tensor copies of a
// stage and an mbarrier counts the bytes in, so the compute warps spend their issue slots on
// LDS + DMMA only.  The hardware zero-fills out-of-range box elements, which removes every
// bounds predicate from the hot loop.  16 warps (4 per scheduler) keep the FP64 tensor pipe fed
// while other warps wait on shared loads or the stage barrier; every warp owns a 16-row x
// 32-column accumulator block:  rank <= 32: 16 row groups (256-row tiles);  rank 33..64:
// 8 row groups x 2 column groups (128-row tiles).
//
//   K-major (contraction mode last):  map dims (k, t, i_n, l), box (16, bt, bi, bl), one box per
//     stage.  Shared layout: rows of 128 B, 16-byte chunk c of row q stored at chunk c ^ (q & 7).
//     Warp rows are visited in the order perm(g) = 2*(g&3) + (g>>2) so that the 16 lanes of a
//     64-bit shared-load phase touch 8 distinct chunks x 2 halves: conflict-free.
//   M-major (contraction mode first): map dims (i_n, l, k), box (16, 1, 16), one box per row group.
//     Shared layout per row group: 16 k-rows x 128 B, swizzled by k & 7; the four k-slots of one
//     DMMA are k = 8j + 2*t + h so the swizzle again spreads them over all banks.
// The factor slice of the contraction mode (16 x R, a few KB) still arrives by cp.async into a
// padded layout; its rows are stored in the matching k order.
// =============================================================================================
constexpr int TK = 16;            // contraction slice per stage of the TMA kernel (one 128-byte swizzle row)
constexpr int TCONS = 512;        // 16 consumer (DMMA) warps
constexpr int TTHREADS = 544;     // + 1 producer warp (TMA + cp.async issue)

template <int NB>
struct TmaCfg {
    static constexpr int CG = (NB > 4) ? 2 : 1;          // column groups of 32
    static constexpr int RG = 16 / CG;                   // row groups of 16
    static constexpr int BMT = 16 * RG;                  // rows per tile
    static constexpr int RS = 8 * NB + 4;                // padded row stride of the factor slice
    static constexpr int XBYTES = BMT * TK * 8;          // 32 KB / 16 KB
    static constexpr int UBYTES = (TK * RS * 8 + 1023) / 1024 * 1024;
    static constexpr int STAGE_BYTES = XBYTES + UBYTES;
    static constexpr int ZC = 32 * CG;                   // staged columns at flush time
    static constexpr int ZSTR = ZC + 2;
    static constexpr int ZBYTES = BMT * ZSTR * 8;
    static constexpr int STAGES = (CG == 1) ? 4 : 6;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + ZBYTES + 512 * 8 + 2 * STAGES * 8 + 1024;
};

// Work schedule shared by the producer and the consumers: CTA b owns items b, b+grid, ...; an item is
// (output-row tile, split) and covers reduction tiles [t0, t1); every tile is nk k-slices.
struct StepState {
    int tau, tau_begin, tau_end;   // linear tile index: tau = it * nred + red  (ntiles < 2^31 on this path)
    int ks;
    __device__ __forceinline__ void init(const DenseParams &p) {
        tau_begin = (int)((p.ntiles * (long long)blockIdx.x) / gridDim.x);
        tau_end = (int)((p.ntiles * (long long)(blockIdx.x + 1)) / gridDim.x);
        tau = tau_begin;
        ks = 0;
    }
    __device__ __forceinline__ bool active(const DenseParams &) const { return tau < tau_end; }
};

template <int NB, bool MMAJOR>
__global__ void __launch_bounds__(TTHREADS, 1)
    mttkrp_dense_tma_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ DenseParams p) {
    using C = TmaCfg<NB>;
    constexpr int STAGES = C::STAGES, RS = C::RS, RG = C::RG;

    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // swizzle atoms need 1 KB alignment
    double *Z = reinterpret_cast<double *>(base + STAGES * C::STAGE_BYTES);
    double *red = reinterpret_cast<double *>(base + STAGES * C::STAGE_BYTES + C::ZBYTES);   // [TCONS] flush scratch
    uint64_t *full = reinterpret_cast<uint64_t *>(red + TCONS);
    uint64_t *empty = full + STAGES;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int bt = 1 << p.lg_bt, bi = 1 << p.lg_bi, bl = 1 << p.lg_bl;

    if (tid == 0) {
        tma_prefetch_desc(&tmX);
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 33);    // 32 cp.async completions (producer lanes) + the TMA expect_tx arrival
            mbar_init(&empty[s], 16);   // one arrival per consumer warp
        }
        fence_barrier_init();
    }
    __syncthreads();

    if (warp == 16) {
        // ================================================================== producer warp
        StepState st;
        st.init(p);
        int stage = 0;
        uint32_t use = 0;   // how many times the current stage slot has been filled before
        while (st.active(p)) {
            if (use > 0) mbar_wait(&empty[stage], (use - 1) & 1u);
            uint8_t *sb = base + stage * C::STAGE_BYTES;
            double *Us = reinterpret_cast<double *>(sb + C::XBYTES);
            const int k0 = st.ks * TK;
            for (int c = lane; c < TK * 4 * NB; c += 32) {
                const int kr = c / (4 * NB);
                const int col = 2 * (c % (4 * NB));
                const int kc = k0 + kr;
                const int dr = MMAJOR ? ((kr & ~7) + 4 * (kr & 1) + ((kr & 7) >> 1)) : kr;
                double *dst = Us + dr * RS + col;
                const double *src = p.Uc + (long long)kc * p.ld + col;
                if (p.vecU) {
                    const bool ok = (kc < p.Kc) && (col < p.R);
                    cp_async16(dst, ok ? src : p.Uc, ok ? 16 : 0);
                } else {
                    const bool ok0 = (kc < p.Kc) && (col < p.R);
                    const bool ok1 = (kc < p.Kc) && (col + 1 < p.R);
                    cp_async8(dst, ok0 ? src : p.Uc, ok0 ? 8 : 0);
                    cp_async8(dst + 1, ok1 ? src + 1 : p.Uc, ok1 ? 8 : 0);
                }
            }
            cp_async_mbar_arrive_noinc(&full[stage]);
            if (lane == 0) {
                const int it = st.tau / p.nred, red = st.tau - it * p.nred;
                const int lt = red / p.ntt, tt = red % p.ntt;
                mbar_arrive_expect_tx(&full[stage], C::XBYTES);
                if (!MMAJOR) {
                    tma_load_4d(sb, &tmX, &full[stage], k0, tt * bt, it * bi, lt * bl);
                } else {
                    // tile = (bi i) x (bl l): one box (16 i, TK k, bl l) per 16-i sub-tile s -- row group rg = s*bl + l_off
                    // holds 16 consecutive i of one l, its TK k-rows of 128 bytes laid out exactly as the consumers expect
                    const int nsub = bi >> 4;
                    for (int sx = 0; sx < nsub; ++sx)
                        tma_load_3d(sb + sx * bl * (TK * 128), &tmX, &full[stage], it * bi + 16 * sx, k0, lt * bl);
                }
            }
            if (++st.ks == p.nk) {
                st.ks = 0;
                ++st.tau;
            }
            if (++stage == STAGES) { stage = 0; ++use; }
        }
        cp_async_wait<0>();
        return;
    }

    // ====================================================================== consumer warps
    const int rg = warp % RG, cg = warp / RG;   // row group / column group of this warp
    const int g = lane >> 2, t4 = lane & 3;
    const int pg = MMAJOR ? g : (2 * (g & 3) + (g >> 2));  // tile row (within a group of 8) owned by fragment row g
    const int nbw = (C::CG == 1) ? (NB < 4 ? NB : 4) : min(4, NB - 4 * cg);  // 8-column blocks this warp computes

    // tile row q -> (l, i, t).  K-major: t fastest, then i, then l.  M-major: row group rg = q >> 4 = s*bl + l_off
    // (16-i sub-tile s, then l), row q & 15 = i within the sub-tile.
    auto decompose = [&](int q, int it, int lt, int tt, int &l, int &i, int &t) -> bool {
        if (MMAJOR) {
            const int rgq = q >> 4;
            t = 0;
            i = it * bi + 16 * (rgq >> p.lg_bl) + (q & 15);
            l = lt * bl + (rgq & (bl - 1));
            return (i < p.In) && (l < p.L);
        }
        const int t_off = q & (bt - 1);
        const int i_off = (q >> p.lg_bt) & (bi - 1);
        const int l_off = q >> (p.lg_bt + p.lg_bi);
        t = tt * bt + t_off;
        i = it * bi + i_off;
        l = lt * bl + l_off;
        return (t < p.T) && (i < p.In) && (l < p.L);
    };
    // tile row that holds (l_off, i_off, t_off) -- the inverse of the above, used by the flush
    auto rowof = [&](int l_off, int i_off, int t_off) -> int {
        if (MMAJOR) return 16 * (((i_off >> 4) << p.lg_bl) + l_off) + (i_off & 15);
        return (((l_off << p.lg_bi) + i_off) << p.lg_bt) + t_off;
    };

    // acc: the running GEMM block of the current tile (registers).  The weighted sums over the tiles of a
    // work item live in shared memory (Z): each thread owns 16 private slots there, so no synchronisation
    // is needed until the item is flushed -- and 32 registers are freed for the DMMA pipeline.
    double acc[2][4][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int nn = 0; nn < 4; ++nn) acc[a][nn][0] = acc[a][nn][1] = 0.0;
    StepState cs;
    cs.init(p);

    // per-thread constant parts of the swizzled A-fragment addresses
    const int a_base = MMAJOR ? (rg * (TK * 128) + (g & 1) * 8) : ((16 * rg + pg) * 128 + (t4 & 1) * 8);
    const int tp = (t4 >> 1) ^ pg;
    const int colbase = 32 * cg;   // first column of this warp's block
    double *zrow0 = Z + (16 * rg + pg) * C::ZSTR + colbase + 2 * t4;   // this thread's slots: rows q, q+8; 4 column pairs
    double *zrow1 = zrow0 + 8 * C::ZSTR;
#pragma unroll
    for (int nn = 0; nn < 4; ++nn) {
        zrow0[8 * nn] = zrow0[8 * nn + 1] = 0.0;
        zrow1[8 * nn] = zrow1[8 * nn + 1] = 0.0;
    }

    int stage = 0;
    uint32_t phase = 0;
    while (cs.active(p)) {
        mbar_wait(&full[stage], phase);
        {
            const uint8_t *Xb = base + stage * C::STAGE_BYTES;
            const double *Us = reinterpret_cast<const double *>(Xb + C::XBYTES) + colbase;
#pragma unroll
            for (int s4 = 0; s4 < TK / 4; ++s4) {
                double a0, a1;
                int brow;
                if (MMAJOR) {
                    const int kk8 = 8 * (s4 >> 1), h = s4 & 1;
                    const int x = 2 * t4 + h;            // == k & 7
                    const int k = kk8 + x;
                    // element (row group rg, k, i_off = 8a + g): chunk (4a + (g>>1)) ^ x
                    a0 = *reinterpret_cast<const double *>(Xb + a_base + k * 128 + (((g >> 1) ^ x) << 4));
                    a1 = *reinterpret_cast<const double *>(Xb + a_base + k * 128 + (((4 + (g >> 1)) ^ x) << 4));
                    brow = kk8 + 4 * h + t4;
                } else {
                    const int kk = 4 * s4;
                    const int off = ((kk >> 1) ^ tp) << 4;
                    a0 = *reinterpret_cast<const double *>(Xb + a_base + off);
                    a1 = *reinterpret_cast<const double *>(Xb + a_base + 8 * 128 + off);
                    brow = kk + t4;
                }
#pragma unroll
                for (int nn = 0; nn < 4; ++nn) {
                    if (nn < nbw) {
                        const double b = Us[brow * RS + 8 * nn + g];
                        dmma884(acc[0][nn][0], acc[0][nn][1], a0, b);
                        dmma884(acc[1][nn][0], acc[1][nn][1], a1, b);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);   // this warp is done reading the slot
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }

        if (++cs.ks == p.nk) {
            cs.ks = 0;
            {
                // ---- tile epilogue: out += acc * prod_m U_m[i_m(row), col]   (the Khatri-Rao rows, on the fly)
                const int it = cs.tau / p.nred, red = cs.tau - it * p.nred;
                const int lt = red / p.ntt, tt = red % p.ntt;
                if (p.direct) {
                    // plain GEMM rows (the partial contraction of the dimension tree, dimtree.cu, and the big TTM of
                    // HOOI): every tile row is its own output row (l, i_n, t) -- no weights, no cross-row sum, 16-byte
                    // stores straight from the accumulator registers, no staging, no barrier
#pragma unroll
                    for (int a = 0; a < 2; ++a) {
                        const int q = 16 * rg + 8 * a + pg;
                        int l, i, t;
                        const bool ok = decompose(q, it, lt, tt, l, i, t);
                        const long long rho = ((long long)l * p.In + i) * p.T + t;
                        double *orow = p.out + rho * p.ld + colbase + 2 * t4;
#pragma unroll
                        for (int nn = 0; nn < 4; ++nn) {
                            const int col = colbase + 8 * nn + 2 * t4;
                            if (nn < nbw && ok) {
                                if (p.direct == 2 && col + 1 < p.R) {
                                    *reinterpret_cast<double2 *>(orow + 8 * nn) = make_double2(acc[a][nn][0], acc[a][nn][1]);
                                } else {
                                    if (col < p.R) orow[8 * nn] = acc[a][nn][0];
                                    if (col + 1 < p.R) orow[8 * nn + 1] = acc[a][nn][1];
                                }
                            }
                            acc[a][nn][0] = acc[a][nn][1] = 0.0;
                        }
                    }
                } else
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    const int q = 16 * rg + 8 * a + pg;
                    int l, i, t;
                    const bool ok = decompose(q, it, lt, tt, l, i, t);
                    double wgt[4][2];
#pragma unroll
                    for (int nn = 0; nn < 4; ++nn) wgt[nn][0] = wgt[nn][1] = ok ? 1.0 : 0.0;
                    if (ok) {
                        for (int m = 0; m < p.nw; ++m) {
                            const int sp = p.w[m].space ? t : l;
                            const int idx = (sp / p.w[m].div) % p.w[m].dim;
                            const double *row = p.w[m].U + (long long)idx * p.ld + colbase;
#pragma unroll
                            for (int nn = 0; nn < 4; ++nn) {
                                const int col = colbase + 8 * nn + 2 * t4;
                                if (col < p.R) wgt[nn][0] *= __ldg(row + 8 * nn + 2 * t4);
                                if (col + 1 < p.R) wgt[nn][1] *= __ldg(row + 8 * nn + 2 * t4 + 1);
                            }
                        }
                    }
                    double *zr = a ? zrow1 : zrow0;
#pragma unroll
                    for (int nn = 0; nn < 4; ++nn) {
                        if (ok) {
                            zr[8 * nn] += wgt[nn][0] * acc[a][nn][0];
                            zr[8 * nn + 1] += wgt[nn][1] * acc[a][nn][1];
                        }
                        acc[a][nn][0] = acc[a][nn][1] = 0.0;
                    }
                }
            }
            const int it_done = cs.tau / p.nred;
            ++cs.tau;
            const bool seg_end = (cs.tau == cs.tau_end) || (cs.tau == (it_done + 1) * p.nred);
            if (seg_end && !p.direct) {
                // ---- segment flush: sum rows of equal i_n in a fixed order, write bi x R outputs.  A segment that covers
                // its output-row tile completely goes straight to M; a partial one goes to this CTA's slab slot (0: the
                // first segment of the range, 1: the last) and is added up, in CTA order, by streamk_fixup_kernel.
                const int it = it_done;
                const int it_lo = it * p.nred, it_hi = it_lo + p.nred;
                const int seg_lo = (cs.tau_begin > it_lo) ? cs.tau_begin : it_lo;
                const bool full = (seg_lo == it_lo) && (cs.tau == it_hi);
                double *dst = full ? (p.out + (long long)it * bi * p.ld)
                                   : (p.slab + (2LL * blockIdx.x + (seg_lo == cs.tau_begin ? 0 : 1)) * p.slab_stride);
                named_bar_sync(1, TCONS);
                const int nout = bi * C::ZC;               // outputs of this item
                const int rows_per_out = bl * bt;          // tile rows that share one output row
                int parts = 1;                             // threads cooperating on one output (power of two)
                while (parts * 2 * nout <= TCONS && parts * 2 <= rows_per_out) parts *= 2;
                if (parts == 1) {
                    for (int o = tid; o < nout; o += TCONS) {
                        const int i_off = o / C::ZC, col = o % C::ZC;
                        const int i = it * bi + i_off;
                        if (col < p.R && i < p.In) {
                            double sum = 0.0;
                            for (int l_off = 0; l_off < bl; ++l_off)
                                for (int t_off = 0; t_off < bt; ++t_off) {
                                    const int q = rowof(l_off, i_off, t_off);
                                    sum += Z[q * C::ZSTR + col];
                                }
                            dst[(long long)i_off * p.ld + col] = sum;
                        }
                    }
                } else {
                    // two-level ordered sum: `parts` threads each add a contiguous run of rows, then one adds the runs
                    const int chunk = rows_per_out / parts;
                    if (tid < nout * parts) {
                        const int part = tid / nout, o = tid % nout;
                        const int i_off = o / C::ZC, col = o % C::ZC;
                        double sum = 0.0;
                        for (int j = part * chunk; j < (part + 1) * chunk; ++j) {
                            const int l_off = j >> p.lg_bt, t_off = j & (bt - 1);
                            const int q = rowof(l_off, i_off, t_off);
                            sum += Z[q * C::ZSTR + col];
                        }
                        red[tid] = sum;
                    }
                    named_bar_sync(1, TCONS);
                    if (tid < nout) {
                        const int i_off = tid / C::ZC, col = tid % C::ZC;
                        const int i = it * bi + i_off;
                        if (col < p.R && i < p.In) {
                            double sum = 0.0;
                            for (int part = 0; part < parts; ++part) sum += red[part * nout + tid];
                            dst[(long long)i_off * p.ld + col] = sum;
                        }
                    }
                }
                named_bar_sync(1, TCONS);
#pragma unroll
                for (int nn = 0; nn < 4; ++nn) {
                    zrow0[8 * nn] = zrow0[8 * nn + 1] = 0.0;
                    zrow1[8 * nn] = zrow1[8 * nn + 1] = 0.0;
                }
            }
        }
    }
}

// M[e] = sum_s P[s][e] in split order (deterministic)
__global__ void reduce_splits_kernel(const double *__restrict__ P, int nsplit, long long n, double *__restrict__ M) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
         e += (long long)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int k = 0; k < nsplit; ++k) s += P[(long long)k * n + e];
        M[e] = s;
    }
}

// Stream-K fix-up: an output-row tile whose reduction tiles were shared by several CTAs gets the sum of their
// partial slabs, in CTA order (deterministic for a fixed grid).  One thread per (it, i_off, column).
__global__ void streamk_fixup_kernel(const double *__restrict__ slab, long long slab_stride, long long ntiles, int nred,
                                     int grid, int nit, int bi, int In, int R, int ld, double *__restrict__ M) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long per_it = (long long)bi * R;
    if (e >= (long long)nit * per_it) return;
    const int it = (int)(e / per_it);
    const int rem = (int)(e - (long long)it * per_it);
    const int i_off = rem / R, col = rem - i_off * R;
    if (it * bi + i_off >= In) return;
    const long long ta = (long long)it * nred, tb = ta + nred - 1;       // first / last tile of this row tile
    // CTA b owns [b*ntiles/grid, (b+1)*ntiles/grid): the owner of tile t is floor(((t+1)*grid - 1) / ntiles)
    const int b_first = (int)((((ta + 1) * grid) - 1) / ntiles);
    const int b_last = (int)((((tb + 1) * grid) - 1) / ntiles);
    if (b_first == b_last) return;                                         // written directly by its single owner
    // grid <= ntiles: no CTA range is empty, so every b in [b_first, b_last] owns a piece of this row tile; only the
    // first one can have started in an earlier row tile (slot 1 = "last segment of its range")
    const long long t0_first = (ntiles * (long long)b_first) / grid;
    const double *src = slab + (long long)i_off * ld + col;
    double sum = src[(2LL * b_first + (t0_first >= ta ? 0 : 1)) * slab_stride];
    int b = b_first + 1;
    for (; b + 4 <= b_last + 1; b += 4) {            // independent loads, added in CTA order
        const double v0 = src[(2LL * b) * slab_stride], v1 = src[(2LL * b + 2) * slab_stride];
        const double v2 = src[(2LL * b + 4) * slab_stride], v3 = src[(2LL * b + 6) * slab_stride];
        sum += v0; sum += v1; sum += v2; sum += v3;
    }
    for (; b <= b_last; ++b) sum += src[(2LL * b) * slab_stride];
    M[((long long)it * bi + i_off) * ld + col] = sum;
}

struct NaiveParams {
    const double *X;
    const double *U[SKT_MAX_NDIM];
    long long shape[SKT_MAX_NDIM];
    long long stride[SKT_MAX_NDIM];
    int ndim, R, mode;
    long long nother;  // prod of the other dims
};

// One thread per output element, serial loop over the rest of the tensor.  Slow and obviously
// correct: the independent on-device check used by the tests.
__global__ void mttkrp_dense_naive_kernel(const __grid_constant__ NaiveParams p, double *__restrict__ M) {
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long In = p.shape[p.mode];
    if (idx >= In * p.R) return;
    const long long i = idx / p.R;
    const int r = (int)(idx % p.R);
    double sum = 0.0;
    for (long long e = 0; e < p.nother; ++e) {
        long long rem = e, off = i * p.stride[p.mode];
        double w = 1.0;
        for (int m = p.ndim - 1; m >= 0; --m) {
            if (m == p.mode) continue;
            const long long im = rem % p.shape[m];
            rem /= p.shape[m];
            off += im * p.stride[m];
            w *= p.U[m][im * p.R + r];
        }
        sum += p.X[off] * w;
    }
    M[idx] = sum;
}

// ------------------------------------------------------------------ host-side planning
struct DensePlan {
    bool mmajor;
    int c;  // contraction mode
    int Kc, In, L, T;
    int lg_bt, lg_bi, lg_bl;
    int nit, nlt, ntt, nred;
    int nsplit, chunk, nitems;
    int nw;
    int wmode[SKT_MAX_NDIM], wdim[SKT_MAX_NDIM], wdiv[SKT_MAX_NDIM], wspace[SKT_MAX_NDIM];
    int grid;
    int bm;  // rows per tile
};

int ilog2(int x) {
    int l = 0;
    while ((1 << l) < x) ++l;
    return l;
}

int make_plan_uncached(const int64_t *shape, int ndim, int rank, int mode, bool tma, DensePlan &pl);

// The plan depends only on (shape, rank, mode, variant); CP-ALS asks for the same N plans every sweep.
int make_plan(const int64_t *shape, int ndim, int rank, int mode, bool tma, DensePlan &pl) {
    struct Key { int64_t shape[SKT_MAX_NDIM]; int ndim, rank, mode, tma, sms; };
    struct Entry { Key k; DensePlan pl; bool used; };
    static thread_local Entry cache[16];
    static thread_local int next = 0;
    if (shape == nullptr || ndim < 2 || ndim > SKT_MAX_NDIM) return make_plan_uncached(shape, ndim, rank, mode, tma, pl);
    Key k;
    memset(&k, 0, sizeof(k));
    for (int m = 0; m < ndim; ++m) k.shape[m] = shape[m];
    k.ndim = ndim; k.rank = rank; k.mode = mode; k.tma = tma ? 1 : 0; k.sms = sm_count();
    for (int e = 0; e < 16; ++e)
        if (cache[e].used && memcmp(&cache[e].k, &k, sizeof(k)) == 0) { pl = cache[e].pl; return SKT_OK; }
    const int rc = make_plan_uncached(shape, ndim, rank, mode, tma, pl);
    if (rc == SKT_OK) {
        cache[next].k = k; cache[next].pl = pl; cache[next].used = true;
        next = (next + 1) % 16;
    }
    return rc;
}

int make_plan_uncached(const int64_t *shape, int ndim, int rank, int mode, bool tma, DensePlan &pl) {
    // rows per tile: 128 for the cp.async kernel; the TMA kernel uses 256 (rank <= 32 per pass) or 128
    const int ncol = std::min(rank, 64);
    const int lgbm = tma ? ((ncol <= 32) ? 8 : 7) : 7;
    const int bm = 1 << lgbm;
    pl.bm = bm;
    SKT_REQUIRE(shape != nullptr, SKT_EINVAL, "shape is NULL");
    SKT_REQUIRE(ndim >= 2 && ndim <= SKT_MAX_NDIM, SKT_EINVAL, "ndim must be in [2, %d], got %d", SKT_MAX_NDIM, ndim);
    SKT_REQUIRE(mode >= 0 && mode < ndim, SKT_EINVAL, "mode %d out of range for a %d-way tensor", mode, ndim);
    SKT_REQUIRE(rank >= 1, SKT_EINVAL, "rank must be >= 1, got %d", rank);
    for (int m = 0; m < ndim; ++m) SKT_REQUIRE(shape[m] >= 1, SKT_ESHAPE, "shape[%d] = %lld", m, (long long)shape[m]);
    pl.mmajor = (mode == ndim - 1);
    pl.c = pl.mmajor ? 0 : ndim - 1;
    long long L = 1, T = 1;
    pl.nw = 0;
    const int lo = pl.mmajor ? 1 : 0;  // first mode of the l-space
    // l-space: modes lo..mode-1 ; t-space: modes mode+1..(c-1) (empty when c == 0)
    {
        long long div = 1;
        for (int m = mode - 1; m >= lo; --m) {
            pl.wmode[pl.nw] = m; pl.wdim[pl.nw] = (int)shape[m]; pl.wdiv[pl.nw] = (int)div; pl.wspace[pl.nw] = 0;
            ++pl.nw;
            div *= shape[m];
            SKT_REQUIRE(div < (1LL << 31), SKT_ESHAPE, "product of the modes left of mode %d exceeds 2^31", mode);
        }
        L = div;
    }
    if (!pl.mmajor) {
        long long div = 1;
        for (int m = ndim - 2; m > mode; --m) {
            pl.wmode[pl.nw] = m; pl.wdim[pl.nw] = (int)shape[m]; pl.wdiv[pl.nw] = (int)div; pl.wspace[pl.nw] = 1;
            ++pl.nw;
            div *= shape[m];
            SKT_REQUIRE(div < (1LL << 31), SKT_ESHAPE, "product of the modes right of mode %d exceeds 2^31", mode);
        }
        T = div;
    }
    SKT_REQUIRE(shape[mode] < (1LL << 31) && shape[pl.c] < (1LL << 31), SKT_ESHAPE, "mode length exceeds 2^31");
    pl.L = (int)L; pl.T = (int)T; pl.In = (int)shape[mode]; pl.Kc = (int)shape[pl.c];

    // tile box: the power-of-two (bl, bi, bt), bl*bi*bt = 128, that wastes the fewest padded rows;
    // ties go to the longest contiguous run (bt for K-major, bi for M-major).
    double best = 1e300;
    int bbt = 0, bbi = 0, bbl = 0;
    for (int lt = 0; lt <= lgbm; ++lt)
        for (int li = 0; li + lt <= lgbm; ++li) {
            const int ll = lgbm - lt - li;
            const long long bt = 1LL << lt, bi = 1LL << li, bl = 1LL << ll;
            if (pl.mmajor && lt != 0) continue;
            if (pl.mmajor && tma && li < 4) continue;   // an M-major TMA row group holds 16 consecutive i
            if (pl.mmajor && tma) {
                static const char *force = getenv("SKT_MM_LI");   // tuning knob: force log2(bi)
                if (force && atoi(force) >= 4 && li != std::min(atoi(force), lgbm)) continue;
            }
            const double padded = (double)((T + bt - 1) / bt * bt) * (double)((pl.In + bi - 1) / bi * bi) *
                                  (double)((L + bl - 1) / bl * bl);
            // prefer (slightly) boxes whose inner run is long: fewer DRAM page switches
            // K-major: long runs in t; M-major (TMA): 32 consecutive i per l measured best (sweep in profiles/r1_notes.md)
            // K-major TMA tiles: 16 t x 16 i (x l) -- measured as fast as long t runs, and it keeps many output-row tiles
            // (few CTAs share one under stream-K: short fix-up, one flush per CTA); the cp.async kernel keeps long t runs
            const int kpref = tma ? (2 * std::abs(lt - 4) + std::abs(li - 4)) : (lgbm - lt);
            const int pref = pl.mmajor ? (tma ? std::abs(li - 5) : (lgbm - li)) : kpref;
            const double score = padded * (1.0 + 1e-3 * pref + 1e-5 * ll);
            if (score < best) { best = score; bbt = lt; bbi = li; bbl = ll; }
        }
    pl.lg_bt = bbt; pl.lg_bi = bbi; pl.lg_bl = bbl;
    const long long nit = (pl.In + (1 << bbi) - 1) >> bbi;
    const long long nlt = (L + (1 << bbl) - 1) >> bbl;
    const long long ntt = (T + (1 << bbt) - 1) >> bbt;
    SKT_REQUIRE(nlt * ntt < (1LL << 31) && nit * nlt * ntt < (1LL << 40), SKT_ESHAPE, "tensor too large for the tile scheduler");
    pl.nit = (int)nit; pl.nlt = (int)nlt; pl.ntt = (int)ntt; pl.nred = (int)(nlt * ntt);

    // split the reduction space so that the static round-robin schedule over the SMs is balanced
    // without drowning the memory system in partial slabs.
    const int G = sm_count();
    const double tile_cost = (double)bm * pl.Kc * G;          // elements streamed, in chip-bandwidth units
    const double slab_cost = 2.0 * (double)pl.In * rank;      // write + re-read of one partial slab
    double best_cost = 1e300;
    int best_ns = 1;
    const int ns_max = (int)std::min<long long>(pl.nred, std::max<long long>(1, (8LL * G + nit - 1) / nit));
    for (int ns = 1; ns <= ns_max; ++ns) {
        const int chunk = (pl.nred + ns - 1) / ns;
        const int ns_eff = (pl.nred + chunk - 1) / chunk;
        if (ns_eff != ns) continue;
        const long long nitems = nit * ns;
        const int last = pl.nred - (ns - 1) * chunk;
        // makespan of the round-robin schedule: CTA b takes items b, b+G, ...
        // items are (it-major, split-minor); the short item of each row-tile is split ns-1.
        long long worst = 0;
        if (nitems <= 4096LL * G) {
            const int gcnt = (int)std::min<long long>(G, nitems);
            for (int b = 0; b < gcnt; ++b) {
                long long tiles = 0;
                for (long long it = b; it < nitems; it += G) tiles += ((it % ns) == ns - 1) ? last : chunk;
                worst = std::max(worst, tiles);
            }
        } else {
            worst = (nitems + G - 1) / G * chunk;
        }
        const double cost = (double)worst * tile_cost + (ns > 1 ? ns * slab_cost : 0.0);
        if (cost < best_cost) { best_cost = cost; best_ns = ns; }
    }
    pl.nsplit = best_ns;
    pl.chunk = (pl.nred + best_ns - 1) / best_ns;
    const long long nitems = nit * best_ns;
    SKT_REQUIRE(nitems < (1LL << 31), SKT_ESHAPE, "too many work items");
    pl.nitems = (int)nitems;
    pl.grid = (int)std::min<long long>(G, nitems);
    return SKT_OK;
}

template <int NB, bool MMAJOR>
int launch_dense(const DenseParams &p, int grid, cudaStream_t st) {
    constexpr int STAGES = (NB <= 4) ? 4 : 3;
    constexpr int RS = 8 * NB + 4;
    constexpr int XSTAGE = MMAJOR ? BK * XS_M : BM * XS_K;
    constexpr int STAGE = XSTAGE + BK * RS;
    const size_t smem = (size_t)(STAGES * STAGE + BM * ZS) * sizeof(double);
    static unsigned long long configured = 0;   /* bit d: configured on device d */  // per template instance
    if (skt::first_use_on_device(configured)) {
        SKT_CUDA(cudaFuncSetAttribute(mttkrp_dense_kernel<NB, MMAJOR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    mttkrp_dense_kernel<NB, MMAJOR><<<grid, NTHREADS, smem, st>>>(p);
    SKT_LAUNCH_CHECK();
    return SKT_OK;
}

// ---- TMA path -------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

bool tma_eligible(const double *X, const DensePlan &pl) {
    const char *off = getenv("SKT_DENSE_NO_TMA");
    if (off && off[0] == '1') return false;
    if ((((uintptr_t)X) & 15) != 0) return false;
    return pl.mmajor ? (pl.In % 2 == 0) : (pl.Kc % 2 == 0);
}

int make_tensor_map(const double *X, const DensePlan &pl, CUtensorMap &tm) {
    EncodeTiledFn enc = encode_tiled_fn();
    SKT_REQUIRE(enc != nullptr, SKT_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[4], strides[3];
    cuuint32_t box[4], estr[4] = {1, 1, 1, 1};
    cuuint32_t rank;
    if (!pl.mmajor) {
        rank = 4;
        dims[0] = pl.Kc; dims[1] = pl.T; dims[2] = pl.In; dims[3] = pl.L;
        strides[0] = (cuuint64_t)pl.Kc * 8;
        strides[1] = strides[0] * pl.T;
        strides[2] = strides[1] * pl.In;
        box[0] = 16; box[1] = 1u << pl.lg_bt; box[2] = 1u << pl.lg_bi; box[3] = 1u << pl.lg_bl;
    } else {
        rank = 3;                                       // (i, k, l): one box = 16 i x TK k x bl l
        dims[0] = pl.In; dims[1] = pl.Kc; dims[2] = pl.L;
        strides[0] = (cuuint64_t)pl.In * 8 * pl.L;       // k
        strides[1] = (cuuint64_t)pl.In * 8;              // l
        box[0] = 16; box[1] = TK; box[2] = 1u << pl.lg_bl;
    }
    const CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, rank, const_cast<double *>(X), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SKT_REQUIRE(r == CUDA_SUCCESS, SKT_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return SKT_OK;
}

template <int NB, bool MMAJOR>
int launch_dense_tma(const CUtensorMap &tm, const DenseParams &p, int grid, cudaStream_t st) {
    const size_t smem = TmaCfg<NB>::SMEM;
    static unsigned long long configured = 0;   /* bit d: configured on device d */
    if (skt::first_use_on_device(configured)) {
        SKT_CUDA(cudaFuncSetAttribute(mttkrp_dense_tma_kernel<NB, MMAJOR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    mttkrp_dense_tma_kernel<NB, MMAJOR><<<grid, TTHREADS, smem, st>>>(tm, p);
    SKT_LAUNCH_CHECK();
    return SKT_OK;
}

template <bool MMAJOR>
int launch_dense_tma_nb(int nb, const CUtensorMap &tm, const DenseParams &p, int grid, cudaStream_t st) {
    switch (nb) {
        case 1: return launch_dense_tma<1, MMAJOR>(tm, p, grid, st);
        case 2: return launch_dense_tma<2, MMAJOR>(tm, p, grid, st);
        case 3: return launch_dense_tma<3, MMAJOR>(tm, p, grid, st);
        case 4: return launch_dense_tma<4, MMAJOR>(tm, p, grid, st);
        case 5: return launch_dense_tma<5, MMAJOR>(tm, p, grid, st);
        case 6: return launch_dense_tma<6, MMAJOR>(tm, p, grid, st);
        case 7: return launch_dense_tma<7, MMAJOR>(tm, p, grid, st);
        default: return launch_dense_tma<8, MMAJOR>(tm, p, grid, st);
    }
}

template <bool MMAJOR>
int launch_dense_nb(int nb, const DenseParams &p, int grid, cudaStream_t st) {
    switch (nb) {
        case 1: return launch_dense<1, MMAJOR>(p, grid, st);
        case 2: return launch_dense<2, MMAJOR>(p, grid, st);
        case 3: return launch_dense<3, MMAJOR>(p, grid, st);
        case 4: return launch_dense<4, MMAJOR>(p, grid, st);
        case 5: return launch_dense<5, MMAJOR>(p, grid, st);
        case 6: return launch_dense<6, MMAJOR>(p, grid, st);
        case 7: return launch_dense<7, MMAJOR>(p, grid, st);
        default: return launch_dense<8, MMAJOR>(p, grid, st);
    }
}

}  // namespace
}  // namespace skt

using namespace skt;

extern "C" size_t skt_mttkrp_dense_workspace(const int64_t *shape, int ndim, int rank, int mode) {
    if (shape && ndim >= 1 && ndim <= SKT_MAX_NDIM)
        for (int m = 0; m < ndim; ++m)
            if (shape[m] == 0) return 256;   // empty tensor: nothing to stage
    // the kernel variant (TMA or cp.async) is chosen at launch from the pointer alignment: cover both plans
    size_t need = 256;
    for (int tma = 0; tma < 2; ++tma) {
        DensePlan pl;
        if (make_plan(shape, ndim, rank, mode, tma != 0, pl) != SKT_OK) return 0;
        if (tma) {   // stream-K slabs: two partial segments of bi x rank per CTA
            const long long g = std::min<long long>(sm_count() + 4, (long long)pl.nit * pl.nred);
            need = std::max(need, align_up((size_t)2 * g * (1 << pl.lg_bi) * rank * sizeof(double), 256));
        } else if (pl.nsplit > 1) {
            need = std::max(need, align_up((size_t)pl.nsplit * pl.In * rank * sizeof(double), 256));
        }
    }
    return need;
}

namespace skt {
// rows_only: store the GEMM rows acc[(l, i_n, t), r] themselves (TTM semantics: no Khatri-Rao weights, no
// reduction over l, t); needs the TMA kernel.  Otherwise the MTTKRP of `mode`.
int dense_impl(const double *X_d, const int64_t *shape, int ndim, const double *const *U_d, int rank, int mode,
               double *M_d, void *ws_d, size_t ws_bytes, skt_stream stream, bool rows_only) {
    // an empty tensor (a rank of a sharded run may own no rows): the MTTKRP is all zeros
    if (shape && ndim >= 1 && ndim <= SKT_MAX_NDIM && mode >= 0 && mode < ndim && rank >= 1) {
        bool empty = false;
        long long rows_out = 1;
        for (int m = 0; m < ndim; ++m) {
            SKT_REQUIRE(shape[m] >= 0, SKT_ESHAPE, "shape[%d] = %lld", m, (long long)shape[m]);
            if (shape[m] == 0) empty = true;
            if (rows_only ? (m != (mode == ndim - 1 ? 0 : ndim - 1)) : (m == mode)) rows_out *= shape[m];
        }
        if (empty) {
            if (rows_out > 0 && M_d) SKT_CUDA(cudaMemsetAsync(M_d, 0, (size_t)rows_out * rank * sizeof(double), (cudaStream_t)stream));
            return SKT_OK;
        }
    }
    SKT_REQUIRE(X_d && U_d && M_d, SKT_EINVAL, "NULL tensor / factor list / output");
    DensePlan pl;
    int rc = make_plan(shape, ndim, rank, mode, false, pl);
    if (rc != SKT_OK) return rc;
    const bool use_tma = tma_eligible(X_d, pl);
    CUtensorMap tm;
    if (use_tma) {
        rc = make_plan(shape, ndim, rank, mode, true, pl);
        if (rc != SKT_OK) return rc;
        rc = make_tensor_map(X_d, pl, tm);
        if (rc != SKT_OK) return rc;
    }
    if (rows_only) {
        SKT_REQUIRE(use_tma, SKT_EUNSUPPORTED, "row-store (TTM) mode needs the TMA kernel (16-byte aligned rows)");
        SKT_REQUIRE(rank <= 64, SKT_EUNSUPPORTED, "row-store (TTM) mode supports at most 64 columns");
        SKT_REQUIRE(U_d[pl.c] != nullptr, SKT_EINVAL, "the matrix of the contracted mode is NULL");
        pl.nw = 0;
    } else {
        for (int m = 0; m < ndim; ++m)
            SKT_REQUIRE(m == mode || U_d[m] != nullptr, SKT_EINVAL, "U[%d] is NULL (only U[mode] may be)", m);
    }
    // every tile row is its own output row: the accumulators are stored straight from registers
    const bool direct = rows_only || (use_tma && !pl.mmajor && pl.L == 1 && pl.T == 1 && pl.nw == 0);
    // TMA kernel: stream-K (contiguous tile ranges, partial segments in per-CTA slabs); cp.async kernel: split slabs
    const bool streamk = use_tma;
    const int nsplit_out = (direct || streamk) ? 1 : pl.nsplit;
    const long long ntiles = (long long)pl.nit * pl.nred;
    SKT_REQUIRE(ntiles < (1LL << 31), SKT_ESHAPE, "tensor too large for the tile scheduler (%lld tiles)", ntiles);
    const int sk_grid = (int)std::min<long long>(sm_count(), ntiles);
    const long long slab_stride = (long long)(1 << pl.lg_bi) * rank;
    const bool sk_partials = streamk && !direct && pl.nred > 1;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t need = sk_partials ? (size_t)2 * sk_grid * slab_stride * sizeof(double)
                                    : ((nsplit_out > 1) ? (size_t)nsplit_out * pl.In * rank * sizeof(double) : 0);
    SKT_REQUIRE(ws_bytes >= need && (need == 0 || ws_d != nullptr), SKT_EWORKSPACE,
                "workspace too small: need %zu bytes, got %zu", need, ws_bytes);

    DenseParams p;
    p.X = X_d;
    p.ld = rank;
    p.Kc = pl.Kc; p.In = pl.In; p.L = pl.L; p.T = pl.T;
    p.NR = (long long)pl.L * pl.In * pl.T;
    p.lg_bt = pl.lg_bt; p.lg_bi = pl.lg_bi; p.lg_bl = pl.lg_bl;
    p.nit = pl.nit; p.ntt = pl.ntt; p.nred = pl.nred;
    p.nsplit = pl.nsplit; p.chunk = pl.chunk; p.nitems = pl.nitems;
    p.nk = use_tma ? (pl.Kc + TK - 1) / TK : (pl.Kc + BK - 1) / BK;
    p.nw = pl.nw;
    p.direct = direct ? 1 : 0;
    p.ntiles = ntiles;
    p.slab = sk_partials ? (double *)ws_d : nullptr;
    p.slab_stride = slab_stride;
    const int grid = streamk ? sk_grid : pl.grid;
    p.out_split_stride = (nsplit_out > 1) ? (long long)pl.In * rank : 0;
    double *outbase = (nsplit_out > 1) ? (double *)ws_d : M_d;
    const bool xalign = (((uintptr_t)X_d) & 15) == 0;
    if (pl.mmajor) p.vecX = xalign && (pl.In % 2 == 0) && (pl.lg_bi >= 1);
    else p.vecX = xalign && (pl.Kc % 2 == 0);

    // rank > 64 runs in column passes of 64 (X is re-streamed once per pass)
    for (int c0 = 0; c0 < rank; c0 += 64) {
        const int ncol = std::min(64, rank - c0);
        p.R = ncol;
        p.Uc = U_d[pl.c] + c0;
        p.out = outbase + c0;
        for (int m = 0; m < pl.nw; ++m) {
            p.w[m].U = U_d[pl.wmode[m]] + c0;
            p.w[m].dim = pl.wdim[m]; p.w[m].div = pl.wdiv[m]; p.w[m].space = pl.wspace[m];
        }
        p.vecU = ((((uintptr_t)p.Uc) & 15) == 0) && (rank % 2 == 0) && (ncol % 2 == 0);
        if (p.direct) p.direct = (((((uintptr_t)p.out) & 15) == 0) && (rank % 2 == 0)) ? 2 : 1;
        const int nb = (ncol + 7) / 8;
        if (use_tma) {
            // the tile height is fixed by the plan: 128-row tiles always run the two-column-group kernels
            const int nbk = (pl.bm == 128) ? std::max(nb, 5) : nb;
            rc = pl.mmajor ? launch_dense_tma_nb<true>(nbk, tm, p, grid, st) : launch_dense_tma_nb<false>(nbk, tm, p, grid, st);
        }
        else
            rc = pl.mmajor ? launch_dense_nb<true>(nb, p, pl.grid, st) : launch_dense_nb<false>(nb, p, pl.grid, st);
        if (rc != SKT_OK) return rc;
        if (sk_partials) {
            const long long n = (long long)pl.nit * (1 << pl.lg_bi) * ncol;
            streamk_fixup_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p.slab, slab_stride, ntiles, pl.nred, grid, pl.nit,
                                                                           1 << pl.lg_bi, pl.In, ncol, rank, p.out);
            SKT_LAUNCH_CHECK();
        }
    }
    if (nsplit_out > 1) {
        const long long n = (long long)pl.In * rank;
        const int blocks = (int)std::min<long long>((n + 255) / 256, 4LL * sm_count());
        reduce_splits_kernel<<<blocks, 256, 0, st>>>((const double *)ws_d, nsplit_out, n, M_d);
        SKT_LAUNCH_CHECK();
    }
    return SKT_OK;
}
}  // namespace skt

extern "C" int skt_mttkrp_dense(const double *X_d, const int64_t *shape, int ndim, const double *const *U_d, int rank,
                                int mode, double *M_d, void *ws_d, size_t ws_bytes, skt_stream stream) {
    return dense_impl(X_d, shape, ndim, U_d, rank, mode, M_d, ws_d, ws_bytes, stream, false);
}

extern "C" int skt_mttkrp_dense_naive(const double *X_d, const int64_t *shape, int ndim, const double *const *U_d,
                                      int rank, int mode, double *M_d, skt_stream stream) {
    SKT_REQUIRE(X_d && shape && U_d && M_d, SKT_EINVAL, "NULL argument");
    SKT_REQUIRE(ndim >= 1 && ndim <= SKT_MAX_NDIM && mode >= 0 && mode < ndim && rank >= 1, SKT_EINVAL, "bad ndim/mode/rank");
    NaiveParams p;
    p.X = X_d; p.ndim = ndim; p.R = rank; p.mode = mode;
    long long stride = 1, nother = 1;
    for (int m = ndim - 1; m >= 0; --m) {
        p.shape[m] = shape[m];
        p.stride[m] = stride;
        stride *= shape[m];
        p.U[m] = U_d[m];
        if (m != mode) nother *= shape[m];
        SKT_REQUIRE(m == mode || U_d[m] != nullptr, SKT_EINVAL, "U[%d] is NULL", m);
    }
    p.nother = nother;
    const long long n = shape[mode] * rank;
    mttkrp_dense_naive_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(p, M_d);
    SKT_LAUNCH_CHECK();
    return SKT_OK;
}

extern "C" int skt_mttkrp_dense_host(const double *X_h, const int64_t *shape, int ndim, const double *const *U_h,
                                     int rank, int mode, double *M_h) {
    SKT_REQUIRE(X_h && shape && U_h && M_h, SKT_EINVAL, "NULL argument");
    SKT_REQUIRE(ndim >= 2 && ndim <= SKT_MAX_NDIM && mode >= 0 && mode < ndim && rank >= 1, SKT_EINVAL, "bad ndim/mode/rank");
    size_t nelem = 1;
    for (int m = 0; m < ndim; ++m) nelem *= (size_t)shape[m];
    double *X_d = nullptr, *M_d = nullptr;
    double *U_d[SKT_MAX_NDIM] = {nullptr};
    void *ws = nullptr;
    int rc = SKT_OK;
    cudaStream_t st = nullptr;
    auto cleanup = [&]() {
        cudaFree(X_d); cudaFree(M_d); cudaFree(ws);
        for (int m = 0; m < ndim; ++m) cudaFree(U_d[m]);
        if (st) cudaStreamDestroy(st);
    };
#define HOST_TRY(call)                                                             \
    do {                                                                           \
        cudaError_t e__ = (call);                                                  \
        if (e__ != cudaSuccess) { cleanup(); return cuda_fail(e__, #call, __FILE__, __LINE__); } \
    } while (0)
    HOST_TRY(cudaStreamCreate(&st));
    HOST_TRY(cudaMalloc(&X_d, nelem * sizeof(double)));
    HOST_TRY(cudaMemcpyAsync(X_d, X_h, nelem * sizeof(double), cudaMemcpyHostToDevice, st));
    for (int m = 0; m < ndim; ++m) {
        if (m == mode) continue;
        if (!U_h[m]) { cleanup(); set_error("U[%d] is NULL", m); return SKT_EINVAL; }
        const size_t b = (size_t)shape[m] * rank * sizeof(double);
        HOST_TRY(cudaMalloc(&U_d[m], b));
        HOST_TRY(cudaMemcpyAsync(U_d[m], U_h[m], b, cudaMemcpyHostToDevice, st));
    }
    const size_t mb = (size_t)shape[mode] * rank * sizeof(double);
    HOST_TRY(cudaMalloc(&M_d, mb));
    const size_t wsb = skt_mttkrp_dense_workspace(shape, ndim, rank, mode);
    HOST_TRY(cudaMalloc(&ws, wsb ? wsb : 256));
    rc = skt_mttkrp_dense(X_d, shape, ndim, U_d, rank, mode, M_d, ws, wsb, st);
    if (rc != SKT_OK) { cleanup(); return rc; }
    HOST_TRY(cudaMemcpyAsync(M_h, M_d, mb, cudaMemcpyDeviceToHost, st));
    HOST_TRY(cudaStreamSynchronize(st));
#undef HOST_TRY
    cleanup();
    return SKT_OK;
}

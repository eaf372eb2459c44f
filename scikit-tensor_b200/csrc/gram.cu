// Unfolding Gram  G = X_(n) X_(n)^T  on the FP64 tensor cores (reference: the Gram inside core.nvecs,
// sktensor/core.py:279,285 -- `Xn = X.unfold(n); Y = Xn.dot(Xn.T)` -- used by the default 'nvecs' initialisation
// of cp.als (cp.py:200-202), by tucker.hosvd (tucker.py:125-137) and inside every HOOI sweep (tucker.py:104)).
//
// The C-order tensor is viewed as (L, I_n, T) and never unfolded:  G[i,j] = sum_{l,t} X[l,i,t] X[l,j,t].
//   K-major (T > 1): both operands of a 128 x 128 output tile are row blocks of the same X_l with the contraction
//                    index t contiguous;  M-major (T == 1, last mode): the rows of X (L x I_n) are the k index and i
//                    is contiguous.
// Only tiles on or above the diagonal are computed (G is symmetric); the (l,t) space is split over grid.y CTAs whose
// partial tiles are added in split order (no atomics: deterministic) and mirrored by the finishing kernel.
// 512 threads = 16 warps, each a 32 x 32 block of DMMA.8x8x4 accumulators; operands staged by cp.async (16-byte) in a
// 3-stage ring.  Shapes whose rows are not 16-byte aligned use the generic strided GEMM in ttm.cu.
#include <algorithm>

#include "common.cuh"

namespace skt {
namespace {

constexpr int GB = 128;          // output tile edge
constexpr int GK = 16;           // contraction slice per stage
constexpr int GTH = 512;
constexpr int GSTAGES = 3;
constexpr int LDK = GK + 4;      // K-major smem row stride (== 4 mod 16: conflict-free fragment loads)
constexpr int LDM = GB + 4;      // M-major smem row stride

struct GramParams {
    const double *X;
    double *out;                 // partial slabs [nsplit][In][In]
    long long L, In, T;
    long long nktiles;           // K-major: L * ceil(T / GK);  M-major: ceil(L / GK)
    long long chunk;             // k-tiles per split
    int ntk;                     // K-major: ceil(T / GK)
    int nb;                      // tiles per edge
};

template <bool MMAJOR>
__global__ void __launch_bounds__(GTH, 1) gram_dmma_kernel(const __grid_constant__ GramParams p) {
    extern __shared__ __align__(16) double gsm[];
    constexpr int TILE = MMAJOR ? GK * LDM : GB * LDK;     // doubles per operand tile
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    // upper-triangular tile pair from blockIdx.x
    int ib = 0, jb = 0;
    {
        int rem = blockIdx.x;
        for (ib = 0; ib < p.nb; ++ib) {
            const int cnt = p.nb - ib;
            if (rem < cnt) { jb = ib + rem; break; }
            rem -= cnt;
        }
    }
    const long long i0 = (long long)ib * GB, j0 = (long long)jb * GB;
    const long long kt0 = (long long)blockIdx.y * p.chunk, kt1 = min(kt0 + p.chunk, p.nktiles);
    const int wr = warp >> 2, wc = warp & 3;               // this warp's 32 x 32 block inside the tile

    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    auto issue = [&](long long kt, int stage) {
        double *As = gsm + (size_t)stage * 2 * TILE, *Bs = As + TILE;
        if (kt < kt1) {
            if (!MMAJOR) {
                const long long l = kt / p.ntk;
                const long long t0 = (kt - l * p.ntk) * GK;
                // 128 rows x 8 chunks of 16 bytes per operand: 2 chunks per thread and operand
#pragma unroll
                for (int rep = 0; rep < 2; ++rep) {
                    const int c = tid + rep * GTH;
                    const int row = c >> 3, ch = c & 7;
                    const long long t = t0 + 2 * ch;
                    const bool okt = t < p.T;           // T is even on this path: a chunk is never split
                    const long long ia = i0 + row, ja = j0 + row;
                    const bool oka = okt && ia < p.In, okb = okt && ja < p.In;
                    cp_async16(As + row * LDK + 2 * ch, p.X + (oka ? ((l * p.In + ia) * p.T + t) : 0), oka ? 16 : 0);
                    cp_async16(Bs + row * LDK + 2 * ch, p.X + (okb ? ((l * p.In + ja) * p.T + t) : 0), okb ? 16 : 0);
                }
            } else {
                const long long l0 = kt * GK;
                // 16 k-rows x 64 chunks per operand
#pragma unroll
                for (int rep = 0; rep < 2; ++rep) {
                    const int c = tid + rep * GTH;
                    const int kr = c >> 6, ch = c & 63;
                    const long long l = l0 + kr;
                    const long long ia = i0 + 2 * ch, ja = j0 + 2 * ch;
                    const bool oka = l < p.L && ia < p.In, okb = l < p.L && ja < p.In;   // In is even on this path
                    cp_async16(As + kr * LDM + 2 * ch, p.X + (oka ? (l * p.In + ia) : 0), oka ? 16 : 0);
                    cp_async16(Bs + kr * LDM + 2 * ch, p.X + (okb ? (l * p.In + ja) : 0), okb ? 16 : 0);
                }
            }
        }
        cp_async_commit();
    };

#pragma unroll
    for (int s = 0; s < GSTAGES - 1; ++s) issue(kt0 + s, s);
    int stage = 0;
    for (long long kt = kt0; kt < kt1; ++kt) {
        cp_async_wait<GSTAGES - 2>();
        __syncthreads();
        issue(kt + GSTAGES - 1, (stage + GSTAGES - 1) % GSTAGES);
        const double *As = gsm + (size_t)stage * 2 * TILE, *Bs = As + TILE;
#pragma unroll
        for (int kk = 0; kk < GK; kk += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int m = 32 * wr + 8 * a + g, n = 32 * wc + 8 * a + g;
                af[a] = MMAJOR ? As[(kk + t4) * LDM + m] : As[m * LDK + kk + t4];
                bf[a] = MMAJOR ? Bs[(kk + t4) * LDM + n] : Bs[n * LDK + kk + t4];
            }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
        }
        stage = (stage + 1) % GSTAGES;
    }
    cp_async_wait<0>();
    // partial tile -> slab of this split
    double *dst = p.out + (size_t)blockIdx.y * p.In * p.In;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const long long i = i0 + 32 * wr + 8 * a + g;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const long long j = j0 + 32 * wc + 8 * b + 2 * t4;
            if (i < p.In) {
                if (j < p.In) dst[i * p.In + j] = acc[a][b][0];
                if (j + 1 < p.In) dst[i * p.In + j + 1] = acc[a][b][1];
            }
        }
    }
}

// G = sum of the splits (in split order); tiles below the diagonal are mirrored from above
__global__ void gram_finish_kernel(const double *__restrict__ P, int nsplit, long long In, double *__restrict__ G) {
    const long long n = In * In;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const long long i = e / In, j = e - i * In;
        const long long src = (i / GB <= j / GB) ? e : (j * In + i);
        double s = 0.0;
        for (int k = 0; k < nsplit; ++k) s += P[(size_t)k * n + src];
        G[e] = s;
    }
}

}  // namespace

// returns SKT_EUNSUPPORTED when the shape / alignment does not fit this kernel (caller falls back)
int gram_dmma(const double *X_d, long long L, long long In, long long T, double *G_d, void *ws_d, size_t ws_bytes,
              cudaStream_t st, size_t *need_out) {
    const bool mmajor = (T == 1);
    const bool ok = (In >= 64) && (mmajor ? (In % 2 == 0) : (T % 2 == 0 && T >= 8)) && ((((uintptr_t)X_d) & 15) == 0 || X_d == nullptr);
    if (!ok) return SKT_EUNSUPPORTED;
    GramParams p;
    p.X = X_d; p.L = L; p.In = In; p.T = T;
    p.ntk = mmajor ? 1 : (int)((T + GK - 1) / GK);
    p.nktiles = mmajor ? (L + GK - 1) / GK : L * p.ntk;
    p.nb = (int)((In + GB - 1) / GB);
    const long long npairs = (long long)p.nb * (p.nb + 1) / 2;
    long long ns = std::max<long long>(1, (2LL * sm_count() + npairs - 1) / npairs);
    ns = std::min<long long>(ns, std::max<long long>(1, p.nktiles / 8));
    ns = std::min<long long>(ns, 512);
    ns = std::min<long long>(ns, std::max<long long>(1, (512LL << 20) / (In * In * 8)));   // at most 512 MB of partial slabs
    p.chunk = (p.nktiles + ns - 1) / ns;
    const int nsplit = (int)((p.nktiles + p.chunk - 1) / p.chunk);
    const size_t need = align_up((size_t)nsplit * In * In * sizeof(double) + 256, 256);
    if (need_out) { *need_out = need; return SKT_OK; }
    SKT_REQUIRE(npairs <= 2147483647LL && nsplit <= 65535, SKT_ESHAPE, "Gram too large for the tile scheduler");
    SKT_REQUIRE(ws_d && ws_bytes >= need, SKT_EWORKSPACE, "workspace too small: need %zu bytes, got %zu", need, ws_bytes);
    p.out = (double *)ws_d;
    const size_t smem = (size_t)GSTAGES * 2 * (mmajor ? GK * LDM : GB * LDK) * sizeof(double);
    dim3 grid((unsigned)npairs, (unsigned)nsplit);
    if (mmajor) {
        static unsigned long long cfg = 0;   /* bit d: configured on device d */
        if (skt::first_use_on_device(cfg)) { SKT_CUDA(cudaFuncSetAttribute(gram_dmma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); }
        gram_dmma_kernel<true><<<grid, GTH, smem, st>>>(p);
    } else {
        static unsigned long long cfg = 0;   /* bit d: configured on device d */
        if (skt::first_use_on_device(cfg)) { SKT_CUDA(cudaFuncSetAttribute(gram_dmma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); }
        gram_dmma_kernel<false><<<grid, GTH, smem, st>>>(p);
    }
    SKT_LAUNCH_CHECK();
    const long long n = In * In;
    const int blocks = (int)std::min<long long>((n + 255) / 256, 4LL * sm_count());
    gram_finish_kernel<<<blocks, 256, 0, st>>>((const double *)ws_d, nsplit, In, G_d);
    SKT_LAUNCH_CHECK();
    return SKT_OK;
}

}  // namespace skt

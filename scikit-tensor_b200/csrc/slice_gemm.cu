// Opt-in: the partial contractions of the dense dimension tree on the INT8 tensor cores (tcgen05.mma.kind::i8, TMEM
// accumulators, TMA), FP64-accurate by error-free slicing.  The default dense path stays the FP64 DMMA kernel of
// mttkrp_dense.cu; this one exists because at rank 32 / 64 that kernel is bound by the 37 TFLOP/s of the FP64 tensor pipe
// (0.28 ms per 512^3 pass) while HBM would allow 0.16 ms.  Prototype and measurements: tools/slice_gemm_probe.cu.
//
//   skt_planes_ttm:   T[row, r] = sum_c X[.., c, ..] U[c, r]   c = the LAST mode (K-major A operand; rows = the other modes in
//                     memory order) or the FIRST mode (MN-major A operand); rank <= 32, contracted mode <= 512
//   skt_planes_pass:  X as the matrix [P x Q]:  T = X KR(U_s..)  or  X^T KR(U_0..U_{s-1})  against the Khatri-Rao product of a
//                     tree group's factors; rank <= 64, any length
//   Both replace skt_mttkrp_dense on the 2-way views of sktensor/_parallel.py (`_tree_pass`), i.e. the unfolded product the
//   reference forms with khatrirao + dot (sktensor/dtensor.py:117-137, core.py:333-378).
//
// Numbers.  v = rint(x * 2^e) is an S-byte two's-complement integer (S = 6: 48 bits, or 5: 40 bits; 2^e is one scale for
// the whole tensor, one per column for the factor).  With 0x80 added to every lower byte its bytes, XOR 0x80, are BALANCED
// digits d_s in [-128, 127], v = sum_s d_s 256^s: one DFMA with the constant 1.5 * 2^52 + 0x80..80 puts them in the low
// mantissa bits, byte permutes move them into S planes.  X is sliced ONCE (it does not change during an ALS run), so a
// pass streams S bytes per element instead of 8.  A slice product sum_c a_i[row, c] b_j[c, r] is an exact INT32 dot
// product (128 * 128 * 6 pairs * 16384 products < 2^31; longer contractions are cut into segments); all B slices
// j >= smin - i that pair with one A slice lie one after the other in shared memory and feed the consecutive accumulators
// i + j, so ONE MMA of N = 32 (S - j0) columns (64 (S - j0) above rank 32, split at 256) covers them.  Pairs with
// i + j < smin (S = 6: 4, S = 5: 2; at most 2^-52 of full scale) are skipped.  The epilogue recombines the accumulators in
// FP64 (Horner in powers of 256) and applies the exact power-of-two scales.
// Error: |dT| <= 2^-(8S-1) max|X| sum_c |U[c, r]| + the same with the roles swapped (normwise 1e-14 .. 1e-13 for S = 6 on
// uniform / normal data, see tests/test_gpu_slices.py); non-finite input poisons the output with NaN.
#include <cuda.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>

#include "common.cuh"

struct skt_planes {
    int ndim;
    int64_t shape[SKT_MAX_NDIM];
    int64_t total;
    int S;
    unsigned char *planes;      // S planes of `total` bytes, each in X's memory order
    double *scale;              // device: [0] * [1] = 2^e, [2] = e, [3] = 0 (NaN when max|X| is not finite)
    unsigned long long *amax;   // device: bits of max|X|
    int device;
    cudaStream_t stream;        // the planes come from the stream-ordered pool of this stream's device
};

namespace skt {
namespace {

constexpr int NB = 32;                  // columns per slice (rank padded with zero digits)
constexpr int TM = 128;                 // rows per tile = TMEM lanes
constexpr int KC = 128;                 // contraction bytes per stage
constexpr int MAXS = 6;
constexpr int MAXK = 512;               // skt_planes_ttm: the factor planes stay resident in shared memory
constexpr int STAGE_BYTES = TM * KC;
constexpr int GEMM_THREADS = 224;        // 7 warps: A producer, MMA issuer, 4 epilogue warps, factor-plane producer

constexpr int MAXACC = 7;               // accumulators (pair sums smin .. 2(S-1)) per set; two sets of 224 columns

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]: 128 x n x 32 bytes of K, signed 8-bit operands, INT32 accumulators (SASS: UTCIMMA)
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
// ---- slicing ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void slice4(const double x[4], double s1, double s2, int S, uint32_t w[MAXS]) {
    uint32_t lo[4], hi[4];
    const double magic = 6755399441055744.0 + (S == 6 ? 551911719040.0 : 2155905152.0);     // 1.5 * 2^52 + 0x80..80
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const double t = fma(x[e] * s1, s2, magic);          // scale = s1 * s2, both powers of two (see split_pow2)
        lo[e] = (uint32_t)__double2loint(t) ^ 0x80808080u;
        hi[e] = (uint32_t)__double2hiint(t) ^ (S == 6 ? 0x80u : 0u);
    }
    const uint32_t t0 = __byte_perm(lo[0], lo[1], 0x5140), t1 = __byte_perm(lo[2], lo[3], 0x5140);
    const uint32_t t2 = __byte_perm(lo[0], lo[1], 0x7362), t3 = __byte_perm(lo[2], lo[3], 0x7362);
    const uint32_t t4 = __byte_perm(hi[0], hi[1], 0x5140), t5 = __byte_perm(hi[2], hi[3], 0x5140);
    w[0] = __byte_perm(t0, t1, 0x5410);
    w[1] = __byte_perm(t0, t1, 0x7632);
    w[2] = __byte_perm(t2, t3, 0x5410);
    w[3] = __byte_perm(t2, t3, 0x7632);
    w[4] = __byte_perm(t4, t5, 0x5410);
    w[5] = __byte_perm(t4, t5, 0x7632);
}

// largest e with rint(amax 2^e) representable by S balanced digits (<= 127 * 256^(S-1)); 0 for amax = 0
__device__ int pow2_exp(double amax, int S) {
    if (!(amax > 0.0)) return 0;
    const int ex = ilogb(amax) + 1;                     // amax = f 2^ex, f in [0.5, 1)
    int e = 8 * S - 1 - ex;
    if (rint(scalbn(amax, e)) > ldexp(127.0, 8 * S - 8)) e -= 1;
    return e;
}
// 2^e as a product of two representable powers of two (e reaches +-1120 for tensors of magnitude 1e-300 / 1e300)
__device__ void split_pow2(int e, double &a, double &b) {
    const int e1 = max(-1000, min(1000, e));
    a = ldexp(1.0, e1);
    b = ldexp(1.0, e - e1);
}

__global__ void planes_absmax_kernel(const double *__restrict__ X, long long n, unsigned long long *out) {
    double m = 0.0;
    bool bad = false;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double a = fabs(X[i]);
        bad |= !(a <= DBL_MAX);
        m = fmax(m, a);
    }
    if (bad) m = __longlong_as_double(0x7ff0000000000000ll);
    for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}

__global__ void planes_scale_kernel(const unsigned long long *amax_bits, int S, double *scale) {
    const double amax = __longlong_as_double((long long)amax_bits[0]);
    if (!(amax <= DBL_MAX)) {
        scale[0] = scale[1] = scale[2] = 0.0;
        scale[3] = nan("");
    } else {
        const int e = pow2_exp(amax, S);
        split_pow2(e, scale[0], scale[1]);
        scale[2] = (double)e;
        scale[3] = 0.0;
    }
}

__global__ void planes_slice_kernel(const double *__restrict__ X, long long total4, long long plane4, const double *scale_d, int S,
                                    uint32_t *__restrict__ P) {
    const double s1 = scale_d[0], s2 = scale_d[1];
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total4; q += (long long)gridDim.x * blockDim.x) {
        const double2 a = reinterpret_cast<const double2 *>(X)[2 * q], b = reinterpret_cast<const double2 *>(X)[2 * q + 1];
        const double x[4] = {a.x, a.y, b.x, b.y};
        uint32_t w[MAXS];
        slice4(x, s1, s2, S, w);
#pragma unroll
        for (int s = 0; s < MAXS; ++s)
            if (s < S) P[s * plane4 + q] = w[s];
    }
}

// factor U [Kc x R] row-major -> planes Bp[s][n][kpad] (K-major rows of U^T, zero beyond Kc and R), one block per column;
// inv[n] = 256^smin / (scale_X * scale_n)
__global__ void __launch_bounds__(128) factor_slice_kernel(const double *__restrict__ U, int Kc, int R, int kpad, int S,
                                                           int smin, const double *__restrict__ xscale,
                                                           uint32_t *__restrict__ Bp, double *__restrict__ inv) {
    __shared__ double red[4];
    __shared__ double sc_s[2];
    const int n = blockIdx.x;
    double m = 0.0;
    bool bad = false;
    if (n < R)
        for (int k = threadIdx.x; k < Kc; k += 128) {
            const double a = fabs(U[(size_t)k * R + n]);
            bad |= !(a <= DBL_MAX);
            m = fmax(m, a);
        }
    if (bad) m = __longlong_as_double(0x7ff0000000000000ll);
    for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        const double amax = fmax(fmax(red[0], red[1]), fmax(red[2], red[3]));
        const bool finite = amax <= DBL_MAX;
        const int e = finite ? pow2_exp(amax, S) : 0;
        sc_s[0] = sc_s[1] = 0.0;
        if (finite) split_pow2(e, sc_s[0], sc_s[1]);
        split_pow2(8 * smin - (int)xscale[2] - e, inv[n], inv[64 + n]);           // 256^smin / (scale_X * scale_n), in two factors
        if (!finite) inv[n] = nan("");
        inv[n] += xscale[3];                                                       // NaN when X is not finite
    }
    __syncthreads();
    const double sc = sc_s[0], sc2 = sc_s[1];
    for (int k4 = threadIdx.x; k4 < kpad / 4; k4 += 128) {
        double x[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int k = 4 * k4 + e;
            x[e] = (n < R && k < Kc) ? U[(size_t)k * R + n] : 0.0;
        }
        uint32_t w[MAXS];
        slice4(x, sc, sc2, S, w);
        const bool pad = n >= R || sc == 0.0;
#pragma unroll
        for (int s = 0; s < MAXS; ++s)
            if (s < S) Bp[((size_t)s * NB + n) * (kpad / 4) + k4] = pad ? 0u : w[s];
    }
}

// ---- the contraction --------------------------------------------------------------------------------------------------
// One kernel for every pass.  A unit of work is (row tile of 128, K segment); a segment is at most 128 chunks of 128 bytes
// (16384 products per INT32 accumulator) and the segments of one tile are summed in order from FP64 slabs.  The factor
// planes are either resident in shared memory (BRES: short contractions, loaded once per CTA) or travel chunk by chunk
// through a two-slot ring beside the A ring.
// Roles: warp 0 = TMA producer of the A ring, warp 6 = TMA producer of the factor planes, warp 1 = MMA issuer (owns the
// tensor memory), warps 2..5 = epilogue (one TMEM lane quarter each).  Warps 0 and 1 run their loops with all 32 lanes and issue through elect.sync: every descriptor is then
// warp-uniform by construction and lives in uniform registers (with a single-lane `if (lane == 0)` loop the compiler wraps
// each UTCIMMA in an election loop and ~35 uniform-datapath instructions, which made the issue thread the bottleneck of
// the rank-64 passes: 1.57 ms per 160^4 pass against 0.9 after this change).
struct GemmParams {
    long long rows;      // output rows (GEMM M)
    int tiles;           // ceil(rows / 128)
    int nchunk;          // ceil(Kc / 128)
    int nseg, cps;       // K segments, chunks per segment
    int Kc;
    int R, NS;
    double *T;           // rows x R (nseg == 1) ...
    double *slabs;       // ... or nseg slabs of rows x R
    const double *inv;   // per column: inv[n] * inv[64 + n] = 256^smin / (scale_X scale_n), two representable powers of two
    // fused second stage (EMODE 2, 3-way tensors): weights W[J x R] of the middle mode, slabs of the finished MTTKRP
    const double *W;
    double *aux;
    int J;               // rows of W
    int arows;           // rows of one aux slab
    int nj, jper, kt;    // EMODE 2: CTAs per 128-row block of the last mode, middle indices per CTA, blocks per fibre
};

// Work of one CTA.  EMODE 0: units (row tile, K segment) round-robin over the grid.  EMODE 2 (first-mode contraction with
// the middle mode folded in): CTA = (128-row block kb of the last mode, range of middle indices j); its units are the tiles
// (j, kb), j ascending, whole contraction each.
template <int EMODE>
struct Units {
    int beg, end, step, kb;
    __device__ __forceinline__ Units(const GemmParams &p) {
        if (EMODE == 2) {
            kb = blockIdx.x / p.nj;
            const int jr = blockIdx.x - kb * p.nj;
            beg = jr * p.jper;
            end = min(p.J, beg + p.jper);
            step = 1;
        } else {
            kb = 0;
            beg = blockIdx.x;
            end = p.tiles * p.nseg;
            step = gridDim.x;
        }
    }
    __device__ __forceinline__ int tile(int u, const GemmParams &p) const { return EMODE == 2 ? u * p.kt + kb : u / p.nseg; }
    __device__ __forceinline__ int seg(int u, const GemmParams &p) const { return EMODE == 2 ? 0 : u - (u / p.nseg) * p.nseg; }
};

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "elect.sync _|P, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, P;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
// one lane polls the barrier, the warp follows (32 lanes polling one mbarrier cost ~5 % of a pass)
__device__ __forceinline__ void mbar_wait_warp(bool elected, uint64_t *bar, uint32_t parity) {
    if (elected) mbar_wait(bar, parity);
    __syncwarp();
}
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);      // SBO 1024 B, descriptor version 1, SWIZZLE_128B
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr) { return ((addr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ uint64_t desc64(uint32_t lo) { return ((uint64_t)DESC_HI << 32) | lo; }
__host__ __device__ constexpr uint32_t idesc_of(int n, bool a_mn) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}
__host__ __device__ constexpr int smin_of(int S) { return S == 6 ? 4 : 2; }       // pairs below it: < 2^-52 of full scale

// blocks [FIRST, FIRST + COUNT) of the stacked B slices, at most 256 columns per instruction
template <bool A_MN, int NBW, int FIRST, int COUNT>
__device__ __forceinline__ void issue_groups(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t acc) {
    constexpr int GMAX = 256 / NBW;
    constexpr uint32_t BT16 = (NBW * KC) >> 4;
#pragma unroll
    for (int b = FIRST; b < FIRST + COUNT; b += GMAX) {
        const int g = (FIRST + COUNT - b) < GMAX ? (FIRST + COUNT - b) : GMAX;
        umma_i8(d + (uint32_t)(b * NBW), desc64(a_lo), desc64(b_lo + (uint32_t)b * BT16), idesc_of(NBW * g, A_MN), acc);
    }
}

// all products of A slice I with the B slices j >= smin - I of one 32-byte k-step.  They lie one after the other in shared
// memory and belong to the consecutive accumulators I + j: one MMA of N = NBW (S - j0) columns (split at 256) covers them.
template <bool A_MN, int NBW, int S, int I>
__device__ __forceinline__ void issue_slice(uint32_t tacc, uint32_t a_lo, uint32_t b_lo_chunk, bool head) {
    constexpr int SMIN = smin_of(S);
    constexpr int J0 = SMIN - I > 0 ? SMIN - I : 0;
    constexpr int NBLK = S - J0;
    constexpr uint32_t BT16 = (NBW * KC) >> 4;
    const uint32_t d = tacc + (uint32_t)((I + J0 - SMIN) * NBW);
    const uint32_t b_lo = b_lo_chunk + (uint32_t)J0 * BT16;
    if (head && I == S - 1) {
        issue_groups<A_MN, NBW, 0, NBLK>(d, a_lo, b_lo, 0u);               // first product of the unit: overwrite
    } else if (head && I >= SMIN) {
        issue_groups<A_MN, NBW, 0, 1>(d, a_lo, b_lo, 0u);                  // opens one more accumulator ...
        issue_groups<A_MN, NBW, 1, NBLK - 1>(d, a_lo, b_lo, 1u);           // ... the others hold data already
    } else {
        issue_groups<A_MN, NBW, 0, NBLK>(d, a_lo, b_lo, 1u);
    }
}

template <bool A_MN, int NBW, int S, int I>
struct SliceLoop {
    // A slices S-1 .. 0 of one chunk: wait for the stage, issue, release it
    static __device__ __forceinline__ void run(const bool elected, const uint32_t tacc, unsigned char *sA, uint64_t *full,
                                               uint64_t *empty, const uint32_t b_lo_chunk, const int nks, const bool first_chunk,
                                               int &st, uint32_t &ph, const int NS) {
        mbar_wait_warp(elected, &full[st], ph);
        tc_fence_after();
        const uint32_t a_lo = desc_lo(smem_u32(sA + (size_t)st * STAGE_BYTES));
        constexpr uint32_t A_STEP16 = (A_MN ? 32 * 128 : 32) >> 4;   // K-major: 32 bytes along the rows; MN-major: 32 rows on
        if (elected) {
#pragma unroll
            for (int ks = 0; ks < KC / 32; ++ks)
                if (ks < nks) issue_slice<A_MN, NBW, S, I>(tacc, a_lo + ks * A_STEP16, b_lo_chunk + ks * 2u, first_chunk && ks == 0);
            umma_commit(&empty[st]);
        }
        __syncwarp();
        if (++st == NS) {
            st = 0;
            ph ^= 1u;
        }
        SliceLoop<A_MN, NBW, S, I - 1>::run(elected, tacc, sA, full, empty, b_lo_chunk, nks, first_chunk, st, ph, NS);
    }
};
template <bool A_MN, int NBW, int S>
struct SliceLoop<A_MN, NBW, S, -1> {
    static __device__ __forceinline__ void run(bool, uint32_t, unsigned char *, uint64_t *, uint64_t *, uint32_t, int, bool, int &,
                                               uint32_t &, int) {}
};

template <bool A_MN, int NBW, int S, bool BRES, int EMODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
slice_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr uint32_t BT = NBW * KC;                              // one B slice tile
    constexpr int NSETS = NBW == 32 ? 2 : 1;                       // accumulator sets in the 512 columns of tensor memory
    constexpr int ACC_COLS = MAXACC * NBW;
    constexpr int SMIN = smin_of(S);
    const int NS = p.NS;
    const int nbslots = BRES ? p.nchunk : 2;
    unsigned char *sB = smem;                                      // nbslots x S tiles
    unsigned char *sA = smem + (size_t)nbslots * S * BT;           // NS stages of 128 x 128 bytes
    uint64_t *bars = (uint64_t *)(sA + (size_t)NS * STAGE_BYTES);
    uint64_t *full = bars, *empty = bars + NS, *bfull = bars + 2 * NS, *bempty = bars + 2 * NS + 2, *tfull = bars + 2 * NS + 4,
             *tempty = bars + 2 * NS + 6;
    uint32_t *tmem_slot = (uint32_t *)(bars + 2 * NS + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bfull[b], 1);
            mbar_init(&bempty[b], 1);
            mbar_init(&tfull[b], 1);
            mbar_init(&tempty[b], 4);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const Units<EMODE> un(p);

    if (warp == 0) {
        // A ring: runs ahead of the MMAs as far as the ring allows, independent of the factor-plane slots
        const bool elected = elect_one();
        int st = 0;
        uint32_t ph = 0;
        bool wrapped = false;
        for (int u = un.beg; u < un.end; u += un.step) {
            const int tile = un.tile(u, p), g = un.seg(u, p);
            const int c0 = g * p.cps, c1 = min(p.nchunk, c0 + p.cps);
            for (int c = c0; c < c1; ++c) {
#pragma unroll
                for (int i = S - 1; i >= 0; --i) {
                    if (wrapped) mbar_wait_warp(elected, &empty[st], ph ^ 1u);
                    if (elected) {
                        mbar_arrive_expect_tx(&full[st], STAGE_BYTES);
                        if (A_MN)
                            tma_load_3d(sA + (size_t)st * STAGE_BYTES, &tmA, &full[st], tile * TM, c * KC, i);
                        else
                            tma_load_3d(sA + (size_t)st * STAGE_BYTES, &tmA, &full[st], c * KC, tile * TM, i);
                    }
                    __syncwarp();
                    if (++st == NS) {
                        st = 0;
                        ph ^= 1u;
                        wrapped = true;
                    }
                }
            }
        }
    } else if (warp == 6) {
        // factor planes: once (resident) or one chunk per slot, two slots
        const bool elected = elect_one();
        if (BRES) {
            if (elected) {
                mbar_arrive_expect_tx(&bfull[0], (uint32_t)(p.nchunk * S * BT));
                for (int c = 0; c < p.nchunk; ++c)
                    for (int j = 0; j < S; ++j) tma_load_3d(sB + (size_t)(c * S + j) * BT, &tmB, &bfull[0], c * KC, 0, j);
            }
            __syncwarp();
        } else {
            int bs = 0;
            uint32_t bph = 0;
            bool bwrapped = false;
            for (int u = un.beg; u < un.end; u += un.step) {
                const int g = un.seg(u, p);
                const int c0 = g * p.cps, c1 = min(p.nchunk, c0 + p.cps);
                for (int c = c0; c < c1; ++c) {
                    if (bwrapped) mbar_wait_warp(elected, &bempty[bs], bph ^ 1u);
                    if (elected) {
                        mbar_arrive_expect_tx(&bfull[bs], (uint32_t)(S * BT));
#pragma unroll
                        for (int j = 0; j < S; ++j) tma_load_3d(sB + ((size_t)bs * S + j) * BT, &tmB, &bfull[bs], c * KC, 0, j);
                    }
                    __syncwarp();
                    if (++bs == 2) {
                        bs = 0;
                        bph ^= 1u;
                        bwrapped = true;
                    }
                }
            }
        }
    } else if (warp == 1) {
        const bool elected = elect_one();
        int st = 0, bs = 0;
        uint32_t ph = 0, bph = 0, lu = 0;
        if (BRES) mbar_wait_warp(elected, &bfull[0], 0);
        for (int u = un.beg; u < un.end; u += un.step, ++lu) {
            const int g = un.seg(u, p);
            const int c0 = g * p.cps, c1 = min(p.nchunk, c0 + p.cps);
            const uint32_t set = lu % NSETS, use = lu / NSETS;
            if (use >= 1) mbar_wait_warp(elected, &tempty[set], (use - 1) & 1u);
            tc_fence_after();
            const uint32_t tacc = tmem + set * (uint32_t)ACC_COLS;
            for (int c = c0; c < c1; ++c) {
                if (!BRES) mbar_wait_warp(elected, &bfull[bs], bph);
                const int valid = min(KC, p.Kc - c * KC);
                const int nks = (valid + 31) >> 5;                     // the zero-filled tail of the last chunk is skipped
                const uint32_t b_lo_chunk = desc_lo(smem_u32(sB + (size_t)(BRES ? c : bs) * S * BT));
                SliceLoop<A_MN, NBW, S, S - 1>::run(elected, tacc, sA, full, empty, b_lo_chunk, nks, c == c0, st, ph, NS);
                if (!BRES) {
                    if (elected) umma_commit(&bempty[bs]);
                    __syncwarp();
                    if (++bs == 2) {
                        bs = 0;
                        bph ^= 1u;
                    }
                }
            }
            if (elected) umma_commit(&tfull[set]);
            __syncwarp();
        }
    } else if (warp < 6) {
        const int q = warp & 3;                                    // the TMEM lane quarter this warp may read
        constexpr int NACC = 2 * (S - 1) - SMIN + 1;
        uint32_t lu = 0;
        double wacc[EMODE == 2 ? 32 : 1];                          // EMODE 2: running sum over the CTA's middle indices
#pragma unroll
        for (int n = 0; n < (EMODE == 2 ? 32 : 1); ++n) wacc[n] = 0.0;
        for (int u = un.beg; u < un.end; u += un.step, ++lu) {
            const int tile = un.tile(u, p), g = un.seg(u, p);
            const uint32_t set = lu % NSETS, use = lu / NSETS;
            const long long row = (long long)tile * TM + q * 32 + lane;
            mbar_wait(&tfull[set], use & 1u);
            tc_fence_after();
            double *out = (p.nseg == 1 ? p.T : p.slabs + (size_t)g * p.rows * p.R) + row * p.R;
#pragma unroll 1
            for (int h = 0; h < NBW / 32; ++h) {
                double r[32];
#pragma unroll
                for (int n = 0; n < 32; ++n) r[n] = 0.0;
#pragma unroll 1
                for (int a = NACC - 1; a >= 0; --a) {
                    uint32_t v[32];
                    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + set * (uint32_t)ACC_COLS + (uint32_t)(a * NBW + h * 32);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                        "tcgen05.wait::ld.sync.aligned;\n"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                        : "r"(taddr)
                        : "memory");
#pragma unroll
                    for (int n = 0; n < 32; ++n) r[n] = fma(r[n], 256.0, (double)(int)v[n]);
                }
                if (EMODE == 2) {
                    // M_last[k, r] += T[(j, k), r] U_mid[j, r]: the tile's rows are 128 consecutive k of fibre j = u
                    const double *w = p.W + (size_t)u * p.R;
#pragma unroll
                    for (int n = 0; n < 32; ++n)
                        if (n < p.R) wacc[n] = fma(r[n] * p.inv[n] * p.inv[64 + n], __ldg(w + n), wacc[n]);
                } else if (row < p.rows) {
                    if (p.R == NBW) {
#pragma unroll
                        for (int n = 0; n < 32; n += 2) {
                            double2 o;
                            o.x = r[n] * p.inv[h * 32 + n] * p.inv[64 + h * 32 + n];
                            o.y = r[n + 1] * p.inv[h * 32 + n + 1] * p.inv[64 + h * 32 + n + 1];
                            *reinterpret_cast<double2 *>(out + h * 32 + n) = o;
                        }
                    } else {
#pragma unroll
                        for (int n = 0; n < 32; ++n)
                            if (h * 32 + n < p.R) out[h * 32 + n] = r[n] * p.inv[h * 32 + n] * p.inv[64 + h * 32 + n];
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[set]);
        }
        if (EMODE == 2) {
            const int jr = blockIdx.x - un.kb * p.nj;
            double *dst = p.aux + ((size_t)jr * p.arows + (size_t)un.kb * TM + q * 32 + lane) * p.R;
#pragma unroll
            for (int n = 0; n < 32; ++n)
                if (n < p.R) dst[n] = wacc[n];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem) : "memory");
}

// ---- general tree pass: long contractions, rank up to 64 --------------------------------------------------------------
// X seen as the matrix [P x Q] (P = product of the modes before the split, Q = of the others):
//   group 0:  T[P x R] = X   KR(U_s .. U_{N-1})      (A operand K-major:  rows p, contraction q)
//   group 1:  T[Q x R] = X^T KR(U_0 .. U_{s-1})      (A operand MN-major: rows q, contraction p)
// KR = the Khatri-Rao product of the group's factors in X's memory order; its rows are formed on the fly by the slicing
// kernel (never stored in FP64).  The factor planes no longer fit shared memory: one k-chunk of them (S tiles of NB x 128
// bytes) travels through a two-slot ring beside the A ring.  A unit of work is (row tile, K segment); a segment is at most
// 128 chunks (16384 products per INT32 accumulator) and segments of one tile are summed in order from FP64 slabs.
struct KrDesc {
    int nfac, R;
    int dims[SKT_MAX_NDIM];
    const double *U[SKT_MAX_NDIM];
};

__device__ __forceinline__ double kr_value(const KrDesc &d, long long k64, int n) {
    double v = 1.0;
    unsigned int k = (unsigned int)k64;                            // the contraction length is below 2^31
    for (int m = d.nfac - 1; m >= 0; --m) {
        const unsigned int idx = k % (unsigned int)d.dims[m];
        k /= (unsigned int)d.dims[m];
        v *= d.U[m][(size_t)idx * d.R + n];
    }
    return v;
}

// column abs-max of KR (bits of a non-negative double order like unsigned integers); 256 threads = 4 rows x 64 columns
__global__ void __launch_bounds__(256) kr_absmax_kernel(const KrDesc d, long long K, unsigned long long *colmax) {
    const int n = threadIdx.x & 63, sub = threadIdx.x >> 6;
    double m = 0.0;
    bool bad = false;
    if (n < d.R)
        for (long long k = (long long)blockIdx.x * 4 + sub; k < K; k += (long long)gridDim.x * 4) {
            const double a = fabs(kr_value(d, k, n));
            bad |= !(a <= DBL_MAX);
            m = fmax(m, a);
        }
    if (bad) m = __longlong_as_double(0x7ff0000000000000ll);
    if (n < d.R && m > 0.0) atomicMax(&colmax[n], (unsigned long long)__double_as_longlong(m));
}

__global__ void kr_scale_kernel(const unsigned long long *colmax, int S, int smin, const double *__restrict__ xscale, double *sc,
                                double *inv) {
    const int n = threadIdx.x;
    const double amax = __longlong_as_double((long long)colmax[n]);
    const bool finite = amax <= DBL_MAX;
    const int e = finite ? pow2_exp(amax, S) : 0;
    sc[n] = sc[64 + n] = 0.0;
    if (finite) split_pow2(e, sc[n], sc[64 + n]);
    split_pow2(8 * smin - (int)xscale[2] - e, inv[n], inv[64 + n]);
    if (!finite) inv[n] = nan("");
    inv[n] += xscale[3];
}

// planes Bp[s][n][kpad] of KR^T; thread = (column n, four consecutive k)
__global__ void __launch_bounds__(256) kr_slice_kernel(const KrDesc d, long long K, long long kpad4, int nb, int S,
                                                       const double *__restrict__ sc, uint32_t *__restrict__ Bp) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= kpad4 * nb) return;
    const int n = (int)(q / kpad4);
    const long long k4 = q % kpad4;
    const double scale = sc[n], scale2 = sc[64 + n];
    double x[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const long long k = 4 * k4 + e;
        x[e] = (n < d.R && k < K) ? kr_value(d, k, n) : 0.0;
    }
    uint32_t w[MAXS];
    slice4(x, scale, scale2, S, w);
    const bool pad = n >= d.R || scale == 0.0;
#pragma unroll
    for (int s = 0; s < MAXS; ++s)
        if (s < S) Bp[((size_t)s * nb + n) * kpad4 + k4] = pad ? 0u : w[s];
}

// T[e] = sum_g slab[g][e], g ascending (fixed order)
__global__ void slab_sum_kernel(const double *__restrict__ slabs, int nseg, long long n, double *__restrict__ T) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        double s = slabs[e];
        for (int g = 1; g < nseg; ++g) s += slabs[(size_t)g * n + e];
        T[e] = s;
    }
}

// ---- host side --------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
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

// byte tensor (inner, outer, slice) with a box of 128 x box_outer x 1, 128-byte swizzle, zero fill outside
int byte_map3(const void *base, uint64_t inner, uint64_t outer, uint64_t slices, uint64_t outer_stride, uint64_t slice_stride,
              uint32_t box_outer, CUtensorMap &tm) {
    EncodeTiledFn enc = encode_fn();
    SKT_REQUIRE(enc != nullptr, SKT_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[3] = {inner, outer, slices}, strides[2] = {outer_stride, slice_stride};
    cuuint32_t box[3] = {KC, box_outer, 1}, estr[3] = {1, 1, 1};
    const CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void *>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SKT_REQUIRE(r == CUDA_SUCCESS, SKT_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return SKT_OK;
}

struct TtmPlan {
    bool a_mn;
    long long rows;
    int Kc, nchunk, kpad;
};

int plan_ttm(const skt_planes *P, int mode, int R, TtmPlan &pl) {
    SKT_REQUIRE(P != nullptr, SKT_EINVAL, "null planes handle");
    SKT_REQUIRE(P->ndim >= 2 && (mode == 0 || mode == P->ndim - 1), SKT_EUNSUPPORTED,
                "the sliced contraction runs over the first or the last mode (got mode %d of %d)", mode, P->ndim);
    SKT_REQUIRE(R >= 1 && R <= NB, SKT_EUNSUPPORTED, "the sliced contraction supports rank 1..%d (got %d)", NB, R);
    pl.a_mn = (mode == 0);
    pl.Kc = (int)P->shape[mode];
    SKT_REQUIRE(pl.Kc >= 1 && pl.Kc <= MAXK, SKT_EUNSUPPORTED, "the sliced contraction supports a contracted mode of 1..%d (got %d)",
                MAXK, pl.Kc);
    pl.rows = P->total / pl.Kc;
    SKT_REQUIRE(P->total % 16 == 0 && (pl.a_mn ? pl.rows % 16 == 0 : pl.Kc % 16 == 0), SKT_EUNSUPPORTED,
                "the sliced contraction needs 16-byte aligned rows of the byte planes");
    SKT_REQUIRE(pl.rows < (1ll << 31) * 1ll, SKT_EUNSUPPORTED, "too many rows");
    pl.nchunk = (pl.Kc + KC - 1) / KC;
    pl.kpad = pl.nchunk * KC;
    return SKT_OK;
}

size_t ttm_ws(const TtmPlan &pl, int S) { return align_up((size_t)S * NB * pl.kpad, 256) + align_up(128 * sizeof(double), 256); }


struct PassPlan {
    bool a_mn;
    long long P, Q, rows, Kc;
    int nfac, fac0;            // the group's factors: modes fac0 .. fac0 + nfac - 1
    int nb;                    // 32 or 64 columns per slice
    int nchunk, nseg, cps;
    long long kpad;
    int tiles;
};

int plan_pass(const skt_planes *P, int split, int group, int R, PassPlan &pl) {
    SKT_REQUIRE(P != nullptr, SKT_EINVAL, "null planes handle");
    SKT_REQUIRE(split >= 1 && split < P->ndim && (group == 0 || group == 1), SKT_EINVAL, "bad split %d / group %d", split, group);
    SKT_REQUIRE(R >= 1 && R <= 64, SKT_EUNSUPPORTED, "the sliced tree pass supports rank 1..64 (got %d)", R);
    pl.P = 1;
    pl.Q = 1;
    for (int m = 0; m < P->ndim; ++m) (m < split ? pl.P : pl.Q) *= P->shape[m];
    SKT_REQUIRE(pl.Q % 16 == 0, SKT_EUNSUPPORTED, "the sliced tree pass needs 16-byte aligned rows of the byte planes");
    pl.a_mn = (group == 1);
    pl.rows = pl.a_mn ? pl.Q : pl.P;
    pl.Kc = pl.a_mn ? pl.P : pl.Q;
    pl.fac0 = pl.a_mn ? 0 : split;
    pl.nfac = pl.a_mn ? split : P->ndim - split;
    pl.nb = R <= 32 ? 32 : 64;
    SKT_REQUIRE(pl.rows < (1ll << 31) && pl.Kc < (1ll << 31), SKT_EUNSUPPORTED, "view too large");
    pl.nchunk = (int)((pl.Kc + KC - 1) / KC);
    pl.kpad = (long long)pl.nchunk * KC;
    pl.tiles = (int)((pl.rows + TM - 1) / TM);
    // K segments: at most 128 chunks each (INT32 range), as many as balance the persistent grid best
    const int sms = sm_count();
    const int min_seg = (pl.nchunk + 127) / 128;
    long long best = -1;
    pl.nseg = min_seg;
    for (int ns = min_seg; ns <= std::max(min_seg, std::min(pl.nchunk, 64)); ++ns) {
        const int cps = (pl.nchunk + ns - 1) / ns;
        const long long waves = ((long long)pl.tiles * ns + sms - 1) / sms;
        const long long cost = waves * (cps + 1);                    // one chunk-time per unit for the epilogue / fill
        if (best < 0 || cost < best) {
            best = cost;
            pl.nseg = ns;
        }
    }
    pl.cps = (pl.nchunk + pl.nseg - 1) / pl.nseg;
    pl.nseg = (pl.nchunk + pl.cps - 1) / pl.cps;
    return SKT_OK;
}

struct PassWs {
    size_t bp, colmax, sc, inv, slabs, total;
};
PassWs pass_ws(const PassPlan &pl, int S, int R) {
    PassWs w;
    size_t off = 0;
    w.bp = off;
    off += align_up((size_t)S * pl.nb * pl.kpad, 256);
    w.colmax = off;
    off += align_up(64 * sizeof(unsigned long long), 256);
    w.sc = off;
    off += align_up(128 * sizeof(double), 256);
    w.inv = off;
    off += align_up(128 * sizeof(double), 256);
    w.slabs = off;
    if (pl.nseg > 1) off += align_up((size_t)pl.nseg * pl.rows * R * sizeof(double), 256);
    w.total = off;
    return w;
}

constexpr size_t BRES_BYTES = 98304;        // factor planes up to this size stay resident in shared memory

template <bool A_MN, int NBW, int S, bool BRES, int EMODE>
int launch_gemm(const CUtensorMap &tmA, const CUtensorMap &tmB, GemmParams &gp, int grid, cudaStream_t st) {
    const size_t bbytes = (size_t)(BRES ? gp.nchunk : 2) * S * NBW * KC;
    const size_t fixed = 1024 + bbytes + 512;                       // alignment slack, factor tiles, barriers
    const size_t cap = smem_optin_bytes();
    SKT_REQUIRE(cap >= fixed + 3 * STAGE_BYTES, SKT_EUNSUPPORTED, "not enough shared memory for the sliced contraction");
    gp.NS = (int)std::min<size_t>(12, (cap - fixed) / STAGE_BYTES);
    const size_t smem = fixed + (size_t)gp.NS * STAGE_BYTES;
    static unsigned long long mask = 0;
    if (first_use_on_device(mask))
        SKT_CUDA(cudaFuncSetAttribute(slice_gemm_kernel<A_MN, NBW, S, BRES, EMODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)cap));
    if (grid <= 0) grid = std::min(gp.tiles * gp.nseg, sm_count());
    slice_gemm_kernel<A_MN, NBW, S, BRES, EMODE><<<grid, GEMM_THREADS, smem, st>>>(tmA, tmB, gp);
    SKT_LAUNCH_CHECK();
    return SKT_OK;
}

template <bool A_MN, int NBW>
int launch_gemm_s(const CUtensorMap &tmA, const CUtensorMap &tmB, GemmParams &gp, int S, cudaStream_t st) {
    const bool bres = (size_t)gp.nchunk * S * NBW * KC <= BRES_BYTES;
    if (S == 6)
        return bres ? launch_gemm<A_MN, NBW, 6, true, 0>(tmA, tmB, gp, 0, st) : launch_gemm<A_MN, NBW, 6, false, 0>(tmA, tmB, gp, 0, st);
    return bres ? launch_gemm<A_MN, NBW, 5, true, 0>(tmA, tmB, gp, 0, st) : launch_gemm<A_MN, NBW, 5, false, 0>(tmA, tmB, gp, 0, st);
}

int launch_gemm_any(bool a_mn, int nb, const CUtensorMap &tmA, const CUtensorMap &tmB, GemmParams &gp, int S, cudaStream_t st) {
    if (a_mn) return nb == 32 ? launch_gemm_s<true, 32>(tmA, tmB, gp, S, st) : launch_gemm_s<true, 64>(tmA, tmB, gp, S, st);
    return nb == 32 ? launch_gemm_s<false, 32>(tmA, tmB, gp, S, st) : launch_gemm_s<false, 64>(tmA, tmB, gp, S, st);
}

}  // namespace
}  // namespace skt

using namespace skt;

extern "C" {

int skt_planes_supported(const int64_t *shape, int ndim, int mode, int rank) {
    if (!shape || ndim < 2 || ndim > SKT_MAX_NDIM || (mode != 0 && mode != ndim - 1) || rank < 1 || rank > NB) return 0;
    long long total = 1;
    for (int m = 0; m < ndim; ++m) {
        if (shape[m] < 1) return 0;
        total *= shape[m];
    }
    const long long Kc = shape[mode], rows = total / Kc;
    if (Kc > MAXK || total % 16 != 0 || rows >= (1ll << 31)) return 0;
    return (mode == 0 ? rows % 16 == 0 : Kc % 16 == 0) ? 1 : 0;
}

int skt_planes_create(const double *X_d, const int64_t *shape, int ndim, int nslices, skt_planes **out, skt_stream stream) {
    cudaStream_t st = (cudaStream_t)stream;
    SKT_REQUIRE(X_d && shape && out, SKT_EINVAL, "null argument");
    SKT_REQUIRE(ndim >= 2 && ndim <= SKT_MAX_NDIM, SKT_EINVAL, "ndim %d outside [2, %d]", ndim, SKT_MAX_NDIM);
    SKT_REQUIRE(nslices == 5 || nslices == 6, SKT_EINVAL, "nslices must be 5 (40-bit) or 6 (48-bit), got %d", nslices);
    int cc_major = 0, dev = 0;
    SKT_CUDA(cudaGetDevice(&dev));
    SKT_CUDA(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    SKT_REQUIRE(cc_major == 10, SKT_EUNSUPPORTED, "the sliced contraction needs the tcgen05 tensor cores of sm_100a");
    long long total = 1;
    for (int m = 0; m < ndim; ++m) {
        SKT_REQUIRE(shape[m] >= 1, SKT_EINVAL, "empty mode %d", m);
        total *= shape[m];
    }
    SKT_REQUIRE(total % 4 == 0 && (((uintptr_t)X_d) & 15) == 0, SKT_EUNSUPPORTED, "the byte planes need a 16-byte aligned tensor");
    skt_planes *P = new skt_planes();
    P->ndim = ndim;
    for (int m = 0; m < ndim; ++m) P->shape[m] = shape[m];
    P->total = total;
    P->S = nslices;
    P->device = dev;
    P->planes = nullptr;
    P->scale = nullptr;
    P->amax = nullptr;
    // stream-ordered pool: a second cp_als call on the same stream gets the block of the first one back without a
    // cudaMalloc (1.5 .. 35 ms for 0.8 GB); the pool keeps up to 8 GB instead of returning them to the driver at once
    P->stream = st;
    static unsigned long long pool_mask = 0;
    if (first_use_on_device(pool_mask)) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = 8ull << 30;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    cudaError_t e = cudaMallocAsync((void **)&P->planes, (size_t)nslices * total, st);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&P->scale, 256, st);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&P->amax, 256, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(P->amax, 0, 8, st);
    if (e != cudaSuccess) {
        skt_planes_destroy(P);
        return cuda_fail(e, "allocating the byte planes", __FILE__, __LINE__);
    }
    const int grid = 8 * sm_count();
    planes_absmax_kernel<<<grid, 256, 0, st>>>(X_d, total, P->amax);
    ++g_launches;
    planes_scale_kernel<<<1, 1, 0, st>>>(P->amax, nslices, P->scale);
    ++g_launches;
    planes_slice_kernel<<<2 * grid, 256, 0, st>>>(X_d, total / 4, total / 4, P->scale, nslices, (uint32_t *)P->planes);
    ++g_launches;
    e = cudaGetLastError();
    if (e != cudaSuccess) {
        skt_planes_destroy(P);
        return cuda_fail(e, "slicing kernels", __FILE__, __LINE__);
    }
    *out = P;
    return SKT_OK;
}

int skt_planes_destroy(skt_planes *P) {
    if (!P) return SKT_OK;
    if (P->planes) cudaFreeAsync(P->planes, P->stream);
    if (P->scale) cudaFreeAsync(P->scale, P->stream);
    if (P->amax) cudaFreeAsync(P->amax, P->stream);
    delete P;
    return SKT_OK;
}

size_t skt_planes_device_bytes(const skt_planes *P) { return P ? (size_t)P->S * P->total + 512 : 0; }
int skt_planes_nslices(const skt_planes *P) { return P ? P->S : 0; }

size_t skt_planes_ttm_workspace(const skt_planes *P, int mode, int rank) {
    TtmPlan pl;
    if (plan_ttm(P, mode, rank, pl) != SKT_OK) return 0;
    return ttm_ws(pl, P->S);
}

int skt_planes_ttm(const skt_planes *P, int mode, const double *U_d, int rank, double *T_d, void *ws_d, size_t ws_bytes,
                   skt_stream stream) {
    cudaStream_t st = (cudaStream_t)stream;
    TtmPlan pl;
    const int rc = plan_ttm(P, mode, rank, pl);
    if (rc) return rc;
    SKT_REQUIRE(U_d && T_d, SKT_EINVAL, "null argument");
    const int S = P->S;
    const size_t need = ttm_ws(pl, S);
    SKT_REQUIRE(ws_d && ws_bytes >= need, SKT_EWORKSPACE, "workspace too small: need %zu bytes, got %zu", need, ws_bytes);
    SKT_REQUIRE((((uintptr_t)ws_d) & 255) == 0 && (((uintptr_t)T_d) & 15) == 0, SKT_EINVAL, "misaligned workspace or output");
    unsigned char *Bp = (unsigned char *)ws_d;
    double *inv = (double *)(Bp + align_up((size_t)S * NB * pl.kpad, 256));

    factor_slice_kernel<<<NB, 128, 0, st>>>(U_d, pl.Kc, rank, pl.kpad, S, smin_of(S), P->scale, (uint32_t *)Bp, inv);
    SKT_LAUNCH_CHECK();

    CUtensorMap tmA, tmB;
    int e;
    if (pl.a_mn)
        e = byte_map3(P->planes, (uint64_t)pl.rows, (uint64_t)pl.Kc, (uint64_t)S, (uint64_t)pl.rows, (uint64_t)P->total, KC, tmA);
    else
        e = byte_map3(P->planes, (uint64_t)pl.Kc, (uint64_t)pl.rows, (uint64_t)S, (uint64_t)pl.Kc, (uint64_t)P->total, TM, tmA);
    if (e) return e;
    e = byte_map3(Bp, (uint64_t)pl.kpad, NB, (uint64_t)S, (uint64_t)pl.kpad, (uint64_t)NB * pl.kpad, NB, tmB);
    if (e) return e;

    GemmParams gp;
    gp.rows = pl.rows;
    gp.tiles = (int)((pl.rows + TM - 1) / TM);
    gp.nchunk = pl.nchunk;
    gp.nseg = 1;
    gp.cps = pl.nchunk;
    gp.Kc = pl.Kc;
    gp.R = rank;
    gp.T = T_d;
    gp.slabs = nullptr;
    gp.inv = inv;
    gp.W = nullptr;
    gp.aux = nullptr;
    gp.J = gp.arows = gp.nj = gp.jper = gp.kt = 0;
    return launch_gemm_any(pl.a_mn, NB, tmA, tmB, gp, S, st);
}

// ---- 3-way tensors: the second stage of the last mode folded into the epilogue of the first-mode contraction ------------
// slabs of M_2[k, r] = sum_j (sum_i X[i, j, k] U[i, r]) W[j, r]; the partial tensor T[(j, k), r] is never written.  One slab of
// shape[2] x rank per CTA range of j; W = the factor of the middle mode.  The slabs are what skt_cp_update_cluster_slabs sums
// while it stages M.  (The same for mode 0 inside the last-mode contraction -- a cross-row reduction by warp shuffles in the
// epilogue -- was measured at +0.06 ms per 512^3 pass against 0.023 ms for the separate reduction kernel and dropped.)
static int fused_plan(const skt_planes *P, int mode, int rank, TtmPlan &pl, int &nslabs, long long &arows, int &nj, int &jper) {
    const int rc = plan_ttm(P, mode, rank, pl);
    if (rc) return rc;
    SKT_REQUIRE(P->ndim == 3 && mode == 0, SKT_EUNSUPPORTED, "the fused second stage is the first-mode contraction of a 3-way tensor");
    SKT_REQUIRE(P->shape[2] % TM == 0, SKT_EUNSUPPORTED, "the fused second stage needs the last mode to be a multiple of 128");
    const long long J = P->shape[1];
    const int nkb = (int)(P->shape[2] / TM);
    nj = (int)std::max<long long>(1, std::min<long long>(J, sm_count() / nkb));
    jper = (int)((J + nj - 1) / nj);
    nj = (int)((J + jper - 1) / jper);
    nslabs = nj;
    arows = P->shape[2];
    return SKT_OK;
}

int skt_planes_fused_slabs(const skt_planes *P, int mode, int rank, int64_t *slab_rows) {
    TtmPlan pl;
    int nslabs = 0, nj, jper;
    long long arows = 0;
    if (fused_plan(P, mode, rank, pl, nslabs, arows, nj, jper) != SKT_OK) return 0;
    if (slab_rows) *slab_rows = arows;
    return nslabs;
}

int skt_planes_ttm_fused(const skt_planes *P, int mode, const double *U_d, const double *W_d, int rank, double *T_d,
                         double *slabs_d, size_t slabs_bytes, void *ws_d, size_t ws_bytes, skt_stream stream) {
    cudaStream_t st = (cudaStream_t)stream;
    TtmPlan pl;
    int nslabs = 0, nj = 0, jper = 0;
    long long arows = 0;
    const int rc = fused_plan(P, mode, rank, pl, nslabs, arows, nj, jper);
    if (rc) return rc;
    SKT_REQUIRE(U_d && W_d && slabs_d, SKT_EINVAL, "null argument");
    SKT_REQUIRE(slabs_bytes >= (size_t)nslabs * arows * rank * sizeof(double), SKT_EWORKSPACE, "slab buffer too small");
    const int S = P->S;
    const size_t need = ttm_ws(pl, S);
    SKT_REQUIRE(ws_d && ws_bytes >= need, SKT_EWORKSPACE, "workspace too small: need %zu bytes, got %zu", need, ws_bytes);
    SKT_REQUIRE((((uintptr_t)ws_d) & 255) == 0 && (((uintptr_t)T_d) & 15) == 0, SKT_EINVAL, "misaligned workspace or output");
    unsigned char *Bp = (unsigned char *)ws_d;
    double *inv = (double *)(Bp + align_up((size_t)S * NB * pl.kpad, 256));
    factor_slice_kernel<<<NB, 128, 0, st>>>(U_d, pl.Kc, rank, pl.kpad, S, smin_of(S), P->scale, (uint32_t *)Bp, inv);
    SKT_LAUNCH_CHECK();
    CUtensorMap tmA, tmB;
    int e;
    if (pl.a_mn)
        e = byte_map3(P->planes, (uint64_t)pl.rows, (uint64_t)pl.Kc, (uint64_t)S, (uint64_t)pl.rows, (uint64_t)P->total, KC, tmA);
    else
        e = byte_map3(P->planes, (uint64_t)pl.Kc, (uint64_t)pl.rows, (uint64_t)S, (uint64_t)pl.Kc, (uint64_t)P->total, TM, tmA);
    if (e) return e;
    e = byte_map3(Bp, (uint64_t)pl.kpad, NB, (uint64_t)S, (uint64_t)pl.kpad, (uint64_t)NB * pl.kpad, NB, tmB);
    if (e) return e;
    GemmParams gp;
    gp.rows = pl.rows;
    gp.tiles = (int)((pl.rows + TM - 1) / TM);
    gp.nchunk = pl.nchunk;
    gp.nseg = 1;
    gp.cps = pl.nchunk;
    gp.Kc = pl.Kc;
    gp.R = rank;
    gp.T = T_d;
    gp.slabs = nullptr;
    gp.inv = inv;
    gp.W = W_d;
    gp.aux = slabs_d;
    gp.J = (int)P->shape[1];
    gp.arows = (int)arows;
    gp.nj = nj;
    gp.jper = jper;
    gp.kt = (int)(P->shape[2] / TM);
    const int grid = gp.kt * nj;
    return S == 6 ? launch_gemm<true, 32, 6, true, 2>(tmA, tmB, gp, grid, st) : launch_gemm<true, 32, 5, true, 2>(tmA, tmB, gp, grid, st);
}

int skt_planes_pass_supported(const int64_t *shape, int ndim, int split, int rank) {
    if (!shape || ndim < 2 || ndim > SKT_MAX_NDIM || split < 1 || split >= ndim || rank < 1 || rank > 64) return 0;
    long long P = 1, Q = 1;
    for (int m = 0; m < ndim; ++m) {
        if (shape[m] < 1) return 0;
        (m < split ? P : Q) *= shape[m];
    }
    return (Q % 16 == 0 && P < (1ll << 31) && Q < (1ll << 31)) ? 1 : 0;
}

size_t skt_planes_pass_workspace(const skt_planes *P, int split, int group, int rank) {
    PassPlan pl;
    if (plan_pass(P, split, group, rank, pl) != SKT_OK) return 0;
    return pass_ws(pl, P->S, rank).total;
}

int skt_planes_pass(const skt_planes *P, int split, int group, const double *const *U_d, int rank, double *T_d, void *ws_d,
                    size_t ws_bytes, skt_stream stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PassPlan pl;
    const int rc = plan_pass(P, split, group, rank, pl);
    if (rc) return rc;
    SKT_REQUIRE(U_d && T_d, SKT_EINVAL, "null argument");
    const int S = P->S;
    const int smin = smin_of(S);
    const PassWs w = pass_ws(pl, S, rank);
    SKT_REQUIRE(ws_d && ws_bytes >= w.total, SKT_EWORKSPACE, "workspace too small: need %zu bytes, got %zu", w.total, ws_bytes);
    SKT_REQUIRE((((uintptr_t)ws_d) & 255) == 0, SKT_EINVAL, "misaligned workspace");
    char *base = (char *)ws_d;
    uint32_t *Bp = (uint32_t *)(base + w.bp);
    unsigned long long *colmax = (unsigned long long *)(base + w.colmax);
    double *sc = (double *)(base + w.sc), *inv = (double *)(base + w.inv), *slabs = (double *)(base + w.slabs);

    KrDesc kd;
    kd.nfac = pl.nfac;
    kd.R = rank;
    for (int f = 0; f < pl.nfac; ++f) {
        kd.dims[f] = (int)P->shape[pl.fac0 + f];
        kd.U[f] = U_d[pl.fac0 + f];
        SKT_REQUIRE(kd.U[f] != nullptr, SKT_EINVAL, "factor %d of the group is null", pl.fac0 + f);
    }
    SKT_CUDA(cudaMemsetAsync(colmax, 0, 64 * sizeof(unsigned long long), st));
    const int kblocks = (int)std::min<long long>((pl.Kc + 3) / 4, 4LL * sm_count());
    kr_absmax_kernel<<<kblocks, 256, 0, st>>>(kd, pl.Kc, colmax);
    SKT_LAUNCH_CHECK();
    kr_scale_kernel<<<1, 64, 0, st>>>(colmax, S, smin, P->scale, sc, inv);
    SKT_LAUNCH_CHECK();
    const long long kpad4 = pl.kpad / 4, nthreads = kpad4 * pl.nb;
    kr_slice_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, st>>>(kd, pl.Kc, kpad4, pl.nb, S, sc, Bp);
    SKT_LAUNCH_CHECK();

    CUtensorMap tmA, tmB;
    int e = byte_map3(P->planes, (uint64_t)pl.Q, (uint64_t)pl.P, (uint64_t)S, (uint64_t)pl.Q, (uint64_t)P->total, 128, tmA);
    if (e) return e;
    e = byte_map3(Bp, (uint64_t)pl.kpad, (uint64_t)pl.nb, (uint64_t)S, (uint64_t)pl.kpad, (uint64_t)pl.nb * pl.kpad, (uint32_t)pl.nb, tmB);
    if (e) return e;

    GemmParams gp;
    gp.rows = pl.rows;
    gp.tiles = pl.tiles;
    gp.nchunk = pl.nchunk;
    gp.nseg = pl.nseg;
    gp.cps = pl.cps;
    gp.Kc = (int)pl.Kc;
    gp.R = rank;
    gp.T = T_d;
    gp.slabs = slabs;
    gp.inv = inv;
    gp.W = nullptr;
    gp.aux = nullptr;
    gp.J = gp.arows = gp.nj = gp.jper = gp.kt = 0;
    e = launch_gemm_any(pl.a_mn, pl.nb, tmA, tmB, gp, S, st);
    if (e) return e;
    if (pl.nseg > 1) {
        const long long n = pl.rows * rank;
        slab_sum_kernel<<<(int)std::min<long long>((n + 255) / 256, 8LL * sm_count()), 256, 0, st>>>(slabs, pl.nseg, n, T_d);
        SKT_LAUNCH_CHECK();
    }
    return SKT_OK;
}

}  // extern "C"

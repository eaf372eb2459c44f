// Prototype: FP64-accurate tall GEMM on the INT8 tensor cores of sm_100a (tcgen05.mma.kind::i8, TMEM accumulators, TMA).
//
//   C[M x 32] = A[M x K] * B[K x 32],   M = 512^2, K = 512   (the dimension-tree pass of BASELINE config 2: the partial
//   contraction T[(i,j), r] = sum_k X[i,j,k] U_2[k,r], DESIGN.md §4.1 "row-store mode")
//
// Error-free slicing with BYTE digits: v = rint(x * 2^e) is held as an S-byte two's-complement integer (S = 5 or 6, the
// scale 2^e global for A, per column for B).  Its bytes ARE the digits: the top byte signed, the others unsigned, so one
// DFMA with the constant 1.5 * 2^52 puts them in the low mantissa bits and byte permutes move them into S planes.
// A slice product  sum_k a_i[m,k] b_j[k,n]  is an exact INT32 dot product on the tensor core; products with the same
// i + j share one TMEM accumulator and the pairs with i + j < SMIN (below 2^-44 of full scale) are skipped.  The epilogue
// recombines the accumulators in FP64 (Horner in powers of 256).
//
// A is sliced ONCE (X does not change during an ALS run): the pass then streams S bytes per element instead of 8.
//
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/bin/slice_gemm_probe
//             tools/slice_gemm_probe.cu -lcuda
// run:   tools/bin/slice_gemm_probe [S=5] [SMIN=3] [dist=0 uniform | 1 normal | 2 lognormal(3)] [logM=18]
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            exit(1);                                                                            \
        }                                                                                       \
    } while (0)

constexpr int K = 512, N = 32, TM = 128, KC = 128, NCHUNK = K / KC;
constexpr int STAGE_BYTES = TM * KC;        // one slice of one k-chunk of one row tile: 128 rows x 128 B, SWIZZLE_128B
constexpr int BTILE_BYTES = N * KC;         // one slice of one k-chunk of B^T: 32 rows x 128 B
constexpr int MAXS = 6;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "W_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra D_%=;\n"
        "bra W_%=;\n"
        "D_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const void *tmap, uint64_t *bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
                     smem_u32(dst)),
                 "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], 128 x 32 x 32 bytes of K, INT8 operands, INT32 accumulate
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
// K-major operand tile, 128-byte rows, SWIZZLE_128B: 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                 // leading byte offset: unused for swizzled K-major
    d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset
    d |= (uint64_t)1 << 46;                 // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
    return d;
}
__host__ __device__ constexpr uint32_t instr_desc(int n) {     // S32 accumulate, signed x signed, both K-major, M = 128
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

struct Params {
    int tiles;           // row tiles of 128
    int S, smin;
    double *C;
    const double *inv_scale_b;      // per column: 2^(8 smin) / (sA * sB[n])
};

// ---------------------------------------------------------------------------------------------------------------------
// slicing: byte planes P[s][row][k] of v = rint(x * scale[row or global])
__device__ __forceinline__ void slice4(const double x[4], double scale, int S, uint32_t w[MAXS]) {
    uint32_t lo[4], hi[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        // 1.5 * 2^52 puts rint(x * scale) into the low mantissa bits (two's complement); the added 0x80 per lower byte and
        // the XOR below turn its bytes into BALANCED digits (every byte signed, -128..127), so that all slice products are
        // signed x signed and the B slices of one A slice can be stacked along N in a single MMA
        const double t = fma(x[e], scale, 6755399441055744.0 + (S == 6 ? 551911719040.0 : 2155905152.0));
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

__global__ void slice_a_kernel(const double *__restrict__ A, long long total4, long long plane, double scale, int S,
                               uint32_t *__restrict__ P) {
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total4; q += (long long)gridDim.x * blockDim.x) {
        const double2 a = reinterpret_cast<const double2 *>(A)[2 * q], b = reinterpret_cast<const double2 *>(A)[2 * q + 1];
        const double x[4] = {a.x, a.y, b.x, b.y};
        uint32_t w[MAXS];
        slice4(x, scale, S, w);
#pragma unroll
        for (int s = 0; s < MAXS; ++s)
            if (s < S) P[s * (plane / 4) + q] = w[s];
    }
}

// B [K x N] row-major -> planes P[s][n][k] (K-major rows of B^T), scale per column
__global__ void slice_b_kernel(const double *__restrict__ B, const double *__restrict__ scale, int S, uint32_t *__restrict__ P) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;       // (n, k/4)
    if (q >= N * K / 4) return;
    const int n = q / (K / 4), k4 = q % (K / 4);
    double x[4];
    for (int e = 0; e < 4; ++e) x[e] = B[(size_t)(4 * k4 + e) * N + n];
    uint32_t w[MAXS];
    slice4(x, scale[n], S, w);
#pragma unroll
    for (int s = 0; s < MAXS; ++s)
        if (s < S) P[(size_t)s * (N * K / 4) + q] = w[s];
}

__global__ void absmax_kernel(const double *__restrict__ A, long long n, unsigned long long *out) {
    double m = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m = fmax(m, fabs(A[i]));
    for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}

// ---------------------------------------------------------------------------------------------------------------------
// main kernel: warp 0 = TMA producer, warp 1 = MMA issuer (owns TMEM), warps 2..5 = epilogue
template <int NS>
__global__ void __launch_bounds__(192, 1)
slice_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Params p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int S = p.S;
    unsigned char *sB = smem;                                   // [chunk][slice] tiles of BTILE_BYTES
    unsigned char *sA = smem + NCHUNK * MAXS * BTILE_BYTES;     // NS stages
    uint64_t *bars = (uint64_t *)(sA + NS * STAGE_BYTES);
    uint64_t *full = bars, *empty = bars + NS, *bfull = bars + 2 * NS, *tfull = bars + 2 * NS + 1, *tempty = bars + 2 * NS + 3;
    uint32_t *tmem_slot = (uint32_t *)(bars + 2 * NS + 5);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(bfull, 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tfull[b], 1);
            mbar_init(&tempty[b], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int nacc = 2 * (S - 1) - p.smin + 1;                  // accumulators: pair sums smin .. 2(S-1)

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(bfull, (uint32_t)(NCHUNK * S * BTILE_BYTES));
            for (int c = 0; c < NCHUNK; ++c)
                for (int j = 0; j < S; ++j) tma_load_2d(sB + (c * MAXS + j) * BTILE_BYTES, &tmB, bfull, c * KC, j * N);
            uint32_t u = 0;
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x)
                for (int c = 0; c < NCHUNK; ++c)
                    for (int i = S - 1; i >= 0; --i, ++u) {
                        const int st = u % NS;
                        if (u >= NS) mbar_wait(&empty[st], ((u / NS) - 1) & 1u);
                        mbar_arrive_expect_tx(&full[st], STAGE_BYTES);
                        tma_load_2d(sA + st * STAGE_BYTES, &tmA, &full[st], c * KC, i * (p.tiles * TM) + tile * TM);
                    }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            mbar_wait(bfull, 0);
            uint32_t u = 0, lt = 0;
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++lt) {
                const uint32_t ab = lt & 1u;
                if (lt >= 2) mbar_wait(&tempty[ab], ((lt >> 1) - 1) & 1u);
                tc_fence_after();
                const uint32_t tacc = tmem + ab * (uint32_t)(MAXS * N);
                for (int c = 0; c < NCHUNK; ++c)
                    for (int i = S - 1; i >= 0; --i, ++u) {
                        const int st = u % NS;
                        mbar_wait(&full[st], (u / NS) & 1u);
                        tc_fence_after();
                        const uint32_t a_addr = smem_u32(sA + st * STAGE_BYTES);
                        // every B slice j >= j0 pairs with this A slice; they lie one after the other in shared memory
                        // (32 rows each) and their products belong to the consecutive accumulators i + j: ONE MMA of
                        // N = 32 (S - j0) columns per 32 bytes of K
                        const int j0 = p.smin - i > 0 ? p.smin - i : 0;
                        const uint32_t b_addr = smem_u32(sB + (c * MAXS + j0) * BTILE_BYTES);
                        const uint32_t idesc = instr_desc(N * (S - j0));
                        const uint32_t d = tacc + (uint32_t)((i + j0 - p.smin) * N);
#pragma unroll
                        for (int ks = 0; ks < KC / 32; ++ks) {
                            const uint64_t da = smem_desc_sw128(a_addr + ks * 32), db = smem_desc_sw128(b_addr + ks * 32);
                            if (c == 0 && ks == 0 && i == S - 1) {
                                umma_i8(d, da, db, idesc, 0u);                  // first product of the tile: overwrite
                            } else if (c == 0 && ks == 0 && i >= p.smin) {
                                // this slice opens one more accumulator (its first 32 columns); the others already hold data
                                umma_i8(d, da, db, instr_desc(N), 0u);
                                umma_i8(d + N, da, smem_desc_sw128(b_addr + BTILE_BYTES + ks * 32), instr_desc(N * (S - 1)), 1u);
                            } else {
                                umma_i8(d, da, db, idesc, 1u);
                            }
                        }
                        umma_commit(&empty[st]);
                    }
                umma_commit(&tfull[ab]);
            }
        }
    } else {
        const int q = warp & 3;                                  // TMEM lane quarter this warp may read
        uint32_t lt = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++lt) {
            const uint32_t ab = lt & 1u;
            mbar_wait(&tfull[ab], (lt >> 1) & 1u);
            tc_fence_after();
            double r[N];
#pragma unroll
            for (int n = 0; n < N; ++n) r[n] = 0.0;
            for (int a = nacc - 1; a >= 0; --a) {
                uint32_t v[N];
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + ab * (uint32_t)(MAXS * N) + (uint32_t)(a * N);
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
                for (int n = 0; n < N; ++n) r[n] = fma(r[n], 256.0, (double)(int)v[n]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[ab]);
            double *out = p.C + ((size_t)tile * TM + q * 32 + lane) * N;
#pragma unroll
            for (int n = 0; n < N; n += 2) {
                double2 o;
                o.x = r[n] * p.inv_scale_b[n];
                o.y = r[n + 1] * p.inv_scale_b[n + 1];
                *reinterpret_cast<double2 *>(out + n) = o;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap byte_map(EncodeTiledFn enc, void *base, uint64_t rows, uint32_t box_rows) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)K, rows}, strides[1] = {(cuuint64_t)K};
    cuuint32_t box[2] = {KC, box_rows}, estr[2] = {1, 1};
    const CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        fprintf(stderr, "cuTensorMapEncodeTiled failed: %d\n", (int)r);
        exit(1);
    }
    return tm;
}

static inline uint64_t splitmix(uint64_t &s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double u01(uint64_t &s) { return (double)(splitmix(s) >> 11) * (1.0 / 9007199254740992.0); }
static double draw(uint64_t &s, int dist) {
    if (dist == 0) return u01(s);
    const double g = sqrt(-2.0 * log(u01(s) + 1e-300)) * cos(6.283185307179586 * u01(s));
    return dist == 1 ? g : exp(3.0 * g);
}
static double pow2_scale(double amax, int S) {          // largest 2^e with amax * 2^e < 2^(8S-1) - 1
    if (!(amax > 0.0)) return 1.0;
    int ex;
    frexp(amax, &ex);                                   // amax = f * 2^ex, f in [0.5, 1)
    double sc = ldexp(1.0, 8 * S - 1 - ex);             // amax * sc in [2^(8S-2), 2^(8S-1))
    if (rint(amax * sc) > ldexp(127.0, 8 * S - 8)) sc *= 0.5;       // balanced digits reach 127.5 * 256^(S-1)
    return sc;
}

int main(int argc, char **argv) {
    const int S = argc > 1 ? atoi(argv[1]) : 5, smin = argc > 2 ? atoi(argv[2]) : 3, dist = argc > 3 ? atoi(argv[3]) : 0;
    const int logM = argc > 4 ? atoi(argv[4]) : 18;
    const long long M = 1ll << logM;
    if (S < 2 || S > MAXS || smin < 0 || 2 * (S - 1) - smin + 1 > MAXS || M % TM) {
        fprintf(stderr, "bad arguments\n");
        return 1;
    }
    printf("C[%lld x %d] = A[%lld x %d] B[%d x %d]; %d byte slices, pair sums >= %d (%d accumulators), dist %d\n", M, N, M, K, K, N,
           S, smin, 2 * (S - 1) - smin + 1, dist);

    std::vector<double> hA((size_t)M * K), hB((size_t)K * N);
    uint64_t seed = 1234567;
    for (auto &x : hA) x = draw(seed, dist);
    for (auto &x : hB) x = draw(seed, dist == 2 ? 1 : dist);

    double *dA, *dB, *dC, *dInv, *dScaleB;
    uint32_t *pA, *pB;
    unsigned long long *dmax;
    CK(cudaMalloc(&dA, (size_t)M * K * 8));
    CK(cudaMalloc(&dB, (size_t)K * N * 8));
    CK(cudaMalloc(&dC, (size_t)M * N * 8));
    CK(cudaMalloc(&dInv, N * 8));
    CK(cudaMalloc(&dScaleB, N * 8));
    CK(cudaMalloc(&pA, (size_t)S * M * K));
    CK(cudaMalloc(&pB, (size_t)S * N * K));
    CK(cudaMalloc(&dmax, 8));
    CK(cudaMemcpy(dA, hA.data(), (size_t)M * K * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), (size_t)K * N * 8, cudaMemcpyHostToDevice));

    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float ms;

    // scales
    CK(cudaMemset(dmax, 0, 8));
    CK(cudaEventRecord(e0));
    absmax_kernel<<<148 * 8, 256>>>(dA, M * K, dmax);
    CK(cudaEventRecord(e1));
    unsigned long long mbits;
    CK(cudaMemcpy(&mbits, dmax, 8, cudaMemcpyDeviceToHost));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double amax;
    memcpy(&amax, &mbits, 8);
    const double sA = pow2_scale(amax, S);
    printf("max|A| = %.6g, scale 2^%d  (abs-max pass %.3f ms)\n", amax, (int)log2(sA), ms);
    std::vector<double> sB(N), inv(N);
    for (int n = 0; n < N; ++n) {
        double m = 0.0;
        for (int k = 0; k < K; ++k) m = fmax(m, fabs(hB[(size_t)k * N + n]));
        sB[n] = pow2_scale(m, S);
        inv[n] = ldexp(1.0, 8 * smin) / (sA * sB[n]);
    }
    CK(cudaMemcpy(dScaleB, sB.data(), N * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dInv, inv.data(), N * 8, cudaMemcpyHostToDevice));

    // slicing (once per tensor / once per factor update)
    CK(cudaEventRecord(e0));
    slice_a_kernel<<<148 * 16, 256>>>(dA, M * K / 4, M * K, sA, S, pA);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("slice A: %.3f ms (%.0f GB/s read+write)\n", ms, (double)M * K * (8 + S) / ms * 1e-6);
    CK(cudaEventRecord(e0));
    slice_b_kernel<<<(N * K / 4 + 255) / 256, 256>>>(dB, dScaleB, S, pB);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("slice B: %.3f ms\n", ms);

    // tensor maps
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres));
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    const CUtensorMap tmA = byte_map(enc, pA, (uint64_t)S * M, TM), tmB = byte_map(enc, pB, (uint64_t)S * N, N);

    constexpr int NS = 8;
    const size_t smem = 1024 + (size_t)NCHUNK * MAXS * BTILE_BYTES + (size_t)NS * STAGE_BYTES + 256;
    CK(cudaFuncSetAttribute(slice_gemm_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    Params p;
    p.tiles = (int)(M / TM);
    p.S = S;
    p.smin = smin;
    p.C = dC;
    p.inv_scale_b = dInv;
    int sms = 148;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int grid = p.tiles < sms ? p.tiles : sms;
    printf("grid %d x 192 threads, %zu bytes of shared memory, %d stages\n", grid, smem, NS);
    CK(cudaMemset(dC, 0xff, (size_t)M * N * 8));
    slice_gemm_kernel<NS><<<grid, 192, smem>>>(tmA, tmB, p);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());

    // check sampled rows against a long-double reference
    std::vector<double> hC((size_t)M * N);
    CK(cudaMemcpy(hC.data(), dC, (size_t)M * N * 8, cudaMemcpyDeviceToHost));
    double worst_norm = 0.0, worst_bound = 0.0, num = 0.0, den = 0.0;
    uint64_t rs = 99;
    const int nsample = 4096;
    for (int t = 0; t < nsample; ++t) {
        const long long m = t < 256 ? (t < 128 ? t : M - 256 + t) : (long long)(splitmix(rs) % (uint64_t)M);
        for (int n = 0; n < N; ++n) {
            long double acc = 0.0L, mag = 0.0L;
            for (int k = 0; k < K; ++k) {
                const long double pr = (long double)hA[(size_t)m * K + k] * (long double)hB[(size_t)k * N + n];
                acc += pr;
                mag += fabsl(pr);
            }
            const double err = fabs((double)((long double)hC[(size_t)m * N + n] - acc));
            num += err * err;
            den += (double)(acc * acc);
            if (mag > 0) worst_bound = fmax(worst_bound, err / (double)mag);
            if (acc != 0) worst_norm = fmax(worst_norm, err / fabs((double)acc));
        }
    }
    printf("sampled %d rows: normwise relative error %.3e, worst |err|/sum|a||b| %.3e, worst elementwise relative %.3e\n", nsample,
           sqrt(num / den), worst_bound, worst_norm);

    // timing
    const int reps = 20;
    for (int w = 0; w < 3; ++w) slice_gemm_kernel<NS><<<grid, 192, smem>>>(tmA, tmB, p);
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) slice_gemm_kernel<NS><<<grid, 192, smem>>>(tmA, tmB, p);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    int npairs = 0;
    for (int i = 0; i < S; ++i)
        for (int j = 0; j < S; ++j) npairs += (i + j >= smin);
    printf("pass: %.4f ms  = %.1f TFLOP/s FP64-equivalent, %.0f GB/s of slice planes, %.0f TOP/s INT8 (%d slice products)\n", ms,
           2.0 * M * K * N / ms * 1e-9, ((double)M * K * S + (double)M * N * 8) / ms * 1e-6, 2.0 * M * K * N * npairs / ms * 1e-9,
           npairs);
    printf("reference points: FP64 DMMA pass of the library 0.281 ms; HBM roof for 8 B/element %.3f ms, for %d B/element %.3f ms\n",
           (double)M * K * 8 / 6541.5e6, S, (double)M * K * S / 6541.5e6);
    return 0;
}

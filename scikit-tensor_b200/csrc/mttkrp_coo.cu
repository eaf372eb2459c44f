// Sparse (COO) FP64 MTTKRP for sm_100a -- the kernel behind sptensor.uttkrp
// (reference: sktensor/sptensor.py:216-229, one ttv / _ttv_compute pass per rank column,
// sptensor.py:153-177: gather-multiply, unique, searchsorted, bincount).
//
// Device format (built once per tensor by skt_coo_build): for every output mode n a copy of the
// non-zeros stably radix-sorted by i_n (own LSD radix sort, radix_sort.cuh) -- int32 subscripts for all N modes plus the FP64 values,
// 4N+8 bytes per non-zero, which is exactly the compulsory index/value traffic of one MTTKRP.
//
// Kernel: all R rank columns in ONE pass.  A group of GS lanes (GS = next power of two >= R,
// at most 32, CPL columns per lane beyond that) walks S consecutive non-zeros; factor rows are
// gathered 8 non-zeros at a time to keep loads in flight, and equal-row runs are accumulated in
// registers.  Complete runs inside a group are stored straight to M; the first/last run of each
// group is merged across the groups of the CTA in shared memory and across CTAs by a short
// fix-up kernel, always in storage order -- a segmented reduction without a single atomic and
// with a reproducible summation order.  Rows without non-zeros are zero (M is cleared first).
#include <algorithm>
#include <ctime>
#include <cstdlib>
#include <type_traits>
#include <cstring>

#include "common.cuh"
#include "radix_sort.cuh"

struct skt_coo {
    int ndim;
    long long nnz;
    long long shape[SKT_MAX_NDIM];
    unsigned built_mask;
    // sorted copy for output mode n: idx[n][m] (nnz int32 each) and vals[n]
    int *idx[SKT_MAX_NDIM][SKT_MAX_NDIM];
    double *vals[SKT_MAX_NDIM];
    // index-panel blocking (see skt_coo_build): the copy for output mode n is sorted by (panel of the largest
    // gathered mode, i_n) and idx[n][n] holds the VIRTUAL row  panel * shape[n] + i_n;  npanel[n] == 1: plain sort
    int npanel[SKT_MAX_NDIM];
    int panel_mode[SKT_MAX_NDIM];
    long long panel_rows[SKT_MAX_NDIM];
    size_t bytes;
};

namespace skt {
namespace {

constexpr int CT = 256;  // threads per CTA
constexpr int S = 256;   // non-zeros per lane group
constexpr int B = 8;     // gather batch (non-zeros in flight per group)

struct CooParams {
    const int *rows;                  // sorted output-mode subscripts
    const int *oidx[SKT_MAX_NDIM];    // subscripts of the other modes (same order)
    const double *U[SKT_MAX_NDIM];    // their factors
    const double *vals;
    long long nnz;
    int R;
    int nother;
    double *M;
    int *carry_row;     // [2*nCTA]
    double *carry_val;  // [2*nCTA][R]
};

// Lane layout.  A group of GS lanes owns S consecutive non-zeros; lane gl of the group owns VPL vectors of
// VEC columns: column (gl + GS*v)*VEC + e.  VEC = 2 (16-byte factor-row loads) whenever R is even.
// All groups of a warp run the same number of batches (validity is a predicate, not a loop bound), so every
// shuffle uses the full warp mask and the compiler is free to hoist the B*NO gathers of a batch above their
// first use -- that is what keeps ~B*NO 128-byte requests per group in flight.
template <int VEC>
struct Vec;
template <>
struct Vec<1> {
    double v[1];
    __device__ __forceinline__ void load(const double *p) { v[0] = __ldg(p); }
};
template <>
struct Vec<2> {
    double v[2];
    __device__ __forceinline__ void load(const double *p) {
        const double2 t = __ldg(reinterpret_cast<const double2 *>(p));
        v[0] = t.x; v[1] = t.y;
    }
};

// STAGED (VEC == 2, VPL == 1): factor rows are gathered with cp.async (LDGSTS) into a per-warp ring of shared-memory
// stages instead of registers.  A lane later reads back exactly the 16-byte chunks it copied itself, so the ring
// needs no barrier -- it is a deep, register-free extension of each lane's load queue: SD-1 batches of
// 8 non-zeros x NO rows x 16 bytes per lane stay in flight (the register version is limited to one batch by its
// 124 registers and 25 % occupancy; ncu: 60 % long-scoreboard stalls with L2 and DRAM both under 40 %).
constexpr int SD = 3;             // ring depth (stages)
// per-group (value,row) block of a stage: SBB x 16 bytes + 16 bytes of padding so that the groups of a warp read
// their broadcast words from different banks
template <int GS, int NO, int SBB>
constexpr size_t staged_warp_bytes() { return (size_t)SD * ((32 / GS) * SBB * NO * GS * 16 + (32 / GS) * (SBB * 16 + 16)); }

template <int GS, int VPL, int VEC, int NO, bool STAGED = false, int SBB = 8, bool CA = false>
__global__ void __launch_bounds__(CT) mttkrp_coo_kernel(const __grid_constant__ CooParams p) {
    constexpr int NG = CT / GS;       // groups per CTA
    constexpr int CPL = VPL * VEC;    // columns per lane
    constexpr int RW = GS * CPL;      // staged row width
    extern __shared__ double csm[];
    int *crow = reinterpret_cast<int *>(csm);                     // [2*NG] per-CTA carry rows: 2 slots per group
    double *cval = csm + ((2 * NG * (int)sizeof(int) + 7) / 8);   // [2*NG][RW]

    const int tid = threadIdx.x;
    const int gl = tid % GS;   // lane within group
    const int gi = tid / GS;   // group within CTA
    const int gbase = (tid & 31) / GS * GS;  // first warp lane of this group
    const long long cta_start = (long long)blockIdx.x * (NG * S);
    const long long start = cta_start + (long long)gi * S;
    const long long end = min(start + (long long)S, p.nnz);
    const int R = p.R;
    constexpr unsigned FULL = 0xffffffffu;
    // (L2 evict_last on the factor rows and evict-first streams were both measured: no gain)

    double acc[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) acc[c] = 0.0;
    int cur_row = -1;
    int nseg = 0;  // completed flushes of this group

    auto colof = [&](int c) { return (gl + GS * (c / VEC)) * VEC + (c % VEC); };

    auto flush = [&](bool is_tail) {
        // first run of the group -> slot 2*gi ; last run (if not also the first) -> slot 2*gi+1 ; others -> M
        if (nseg == 0 || is_tail) {
            const int slot = 2 * gi + (nseg == 0 ? 0 : 1);
            if (gl == 0) crow[slot] = cur_row;
#pragma unroll
            for (int c = 0; c < CPL; ++c) cval[slot * RW + colof(c)] = acc[c];
        } else {
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                const int col = colof(c);
                if (col < R) p.M[(long long)cur_row * R + col] = acc[c];
            }
        }
        ++nseg;
    };

    if (gl == 0) { crow[2 * gi] = -1; crow[2 * gi + 1] = -1; }

    if (start < end) cur_row = p.rows[start];
    // uniform trip count for the whole CTA: S / BB batches (the last groups of the tensor run on predicates)
    constexpr int BB = (NO <= 3 && VPL <= 2) ? B : 2;   // gather batch; the many-mode and wide-rank variants keep fewer rows in registers
    int ccol[VPL];                          // first column of each of this lane's vectors, clamped into the row
    bool cok[VPL];
#pragma unroll
    for (int vv = 0; vv < VPL; ++vv) {
        const int col0 = (gl + GS * vv) * VEC;
        cok[vv] = col0 < R;
        ccol[vv] = cok[vv] ? col0 : 0;
    }
    if constexpr (STAGED) {
        static_assert(!STAGED || (VEC == 2 && VPL == 1), "staged gathers use one 16-byte chunk per lane and row");
        constexpr int NGW = 32 / GS;                       // groups per warp
        constexpr int ROWB = GS * 16;                      // bytes per staged row
        constexpr int DATAB = NGW * SBB * NO * ROWB;       // row data per stage
        constexpr int METAG = SBB * 16 + 16;               // padded (value, row id) block of one group
        constexpr int STAGEB = DATAB + NGW * METAG;        // + (value, row id) per non-zero
        constexpr int NBATCH = S / SBB;                    // 32 batches per group
        constexpr int RND = 4;                             // batches per subscript round
        unsigned char *ring = reinterpret_cast<unsigned char *>(cval + 2 * NG * RW) + (size_t)(tid >> 5) * (SD * STAGEB);
        const int gw = (tid & 31) / GS;                    // group within the warp
        const bool colok = cok[0];
        // subscripts / values of one round (4 batches x 8 non-zeros per group): lane gl < 8 holds non-zero 8c + gl
        int r_row[2][RND], r_idx[2][RND][NO];
        double r_val[2][RND];
        // (round buffers and components are compile-time constants everywhere: the arrays stay in registers)
        auto load_round = [&](int rnd, auto BUF) {
            constexpr int buf = decltype(BUF)::value;
#pragma unroll
            for (int c = 0; c < RND; ++c) {
                const long long q = start + (long long)rnd * (RND * SBB) + c * SBB + gl;
                const bool have = (gl < SBB) && (q < end) && (rnd * RND + c < NBATCH);
                // (L2 evict_last on the factor rows / streaming loads here were measured: no effect at 200 M nnz)
                r_row[buf][c] = have ? p.rows[q] : -1;
                r_val[buf][c] = have ? p.vals[q] : 0.0;
#pragma unroll
                for (int m = 0; m < NO; ++m) r_idx[buf][c][m] = (have && m < p.nother) ? p.oidx[m][q] : 0;
            }
        };
        auto issue_batch = [&](int k, auto BUF, auto COMP) {    // batch k lives in round buffer BUF, component COMP
            constexpr int buf = decltype(BUF)::value, c = decltype(COMP)::value;
            unsigned char *st = ring + (k % SD) * STAGEB;
#pragma unroll
            for (int b = 0; b < SBB; ++b) {
                const int src = gbase + b;
#pragma unroll
                for (int m = 0; m < NO; ++m) {
                    const int im = __shfl_sync(FULL, r_idx[buf][c][m], src);
                    const int mm = (NO <= 3 || m < p.nother) ? m : 0;
                    const double *rowp = p.U[mm] + (long long)((NO <= 3 || m < p.nother) ? im : 0) * R + ccol[0];
                    if constexpr (CA) cp_async16_ca(st + ((gw * SBB + b) * NO + m) * ROWB + gl * 16, rowp);
                    else cp_async16(st + ((gw * SBB + b) * NO + m) * ROWB + gl * 16, rowp, 16);
                }
            }
            if (gl < SBB) {
                double *meta = reinterpret_cast<double *>(st + DATAB + gw * METAG + gl * 16);
                *reinterpret_cast<double2 *>(meta) = make_double2(r_val[buf][c], __hiloint2double(0, r_row[buf][c]));
            }
        };
        auto compute_batch = [&](int t) {
            const unsigned char *st = ring + (t % SD) * STAGEB;
#pragma unroll
            for (int b = 0; b < SBB; ++b) {
                // (value, row id) of the non-zero in one 16-byte broadcast load
                const double2 mt = *reinterpret_cast<const double2 *>(st + DATAB + gw * METAG + b * 16);
                const double v = mt.x;
                const int row = __double2loint(mt.y);
                double x0 = colok ? v : 0.0, x1 = x0;
#pragma unroll
                for (int m = 0; m < NO; ++m) {
                    if (NO <= 3 || m < p.nother) {
                        const double2 u = *reinterpret_cast<const double2 *>(st + ((gw * SBB + b) * NO + m) * ROWB + gl * 16);
                        x0 *= u.x; x1 *= u.y;
                    }
                }
                if (row >= 0) {
                    if (row != cur_row) {
                        flush(false);
                        acc[0] = acc[1] = 0.0;
                        cur_row = row;
                    }
                    acc[0] += x0; acc[1] += x1;
                }
            }
        };
        using I0 = std::integral_constant<int, 0>;
        using I1 = std::integral_constant<int, 1>;
        // one round: RND batches whose subscripts sit in buffer CUR; the ring is refilled SD-1 batches ahead, from
        // this round's registers or (for the last SD-1 batches) from the next round's
        auto round_body = [&](int rnd, auto CUR) {
            constexpr int cur = decltype(CUR)::value;
            using C = std::integral_constant<int, cur>;
            using N = std::integral_constant<int, cur ^ 1>;
            auto step = [&](auto CC) {
                constexpr int c = decltype(CC)::value;
                const int t = rnd * RND + c;
                const int k = t + SD - 1;
                constexpr int kc = (c + SD - 1) % RND;
                if (k < NBATCH) {
                    if constexpr (c + SD - 1 >= RND) issue_batch(k, N{}, std::integral_constant<int, kc>{});
                    else issue_batch(k, C{}, std::integral_constant<int, kc>{});
                }
                cp_async_commit();
                cp_async_wait<SD - 1>();
                __syncwarp();                              // the (value, row) pairs were written by other lanes
                compute_batch(t);
                __syncwarp();                              // stage t % SD is refilled by the next issue
            };
            step(std::integral_constant<int, 0>{});
            step(std::integral_constant<int, 1>{});
            step(std::integral_constant<int, 2>{});
            step(std::integral_constant<int, 3>{});
            load_round(rnd + 2, C{});                      // this round's registers are free: fetch round rnd + 2
        };
        load_round(0, I0{});
        load_round(1, I1{});
        issue_batch(0, I0{}, I0{});
        cp_async_commit();
        issue_batch(1, I0{}, I1{});
        cp_async_commit();
        static_assert(SD == 3 && RND == 4 && (NBATCH / RND) % 2 == 0, "prologue / round pairing written for SD = 3");
        for (int rnd = 0; rnd < NBATCH / RND; rnd += 2) {
            round_body(rnd, I0{});
            round_body(rnd + 1, I1{});
        }
        cp_async_wait<0>();
    } else
    for (int b0 = 0; b0 < S; b0 += BB) {
        // lanes 0..BB-1 of the group fetch the subscripts/value of one non-zero each (BB <= GS)
        const long long my = start + b0 + gl;
        const bool have = (gl < BB) && (my < end);
        const int my_row = have ? p.rows[my] : -1;
        const double my_val = have ? p.vals[my] : 0.0;
        int my_idx[NO];
#pragma unroll
        for (int m = 0; m < NO; ++m) my_idx[m] = (have && m < p.nother) ? p.oidx[m][my] : 0;

        int rws[BB];
        double vs[BB];
        Vec<VEC> u[BB][NO][VPL];
        // phase 1: every gather of the batch is issued before any is used (addresses are always valid:
        // padding entries carry subscript 0, lanes beyond R read column 0 and are masked below)
#pragma unroll
        for (int b = 0; b < BB; ++b) {
            const int src = gbase + b;
            rws[b] = __shfl_sync(FULL, my_row, src);
            vs[b] = __shfl_sync(FULL, my_val, src);
#pragma unroll
            for (int m = 0; m < NO; ++m) {
                const int im = __shfl_sync(FULL, my_idx[m], src);
                const int mm = (NO <= 3 || m < p.nother) ? m : 0;
                const double *rowp = p.U[mm] + (long long)((NO <= 3 || m < p.nother) ? im : 0) * R;
#pragma unroll
                for (int vv = 0; vv < VPL; ++vv) u[b][m][vv].load(rowp + ccol[vv]);
            }
        }
        // phase 2: products, then the ordered run accumulation
#pragma unroll
        for (int b = 0; b < BB; ++b) {
            double contrib[CPL];
#pragma unroll
            for (int vv = 0; vv < VPL; ++vv)
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    double x = cok[vv] ? vs[b] : 0.0;
#pragma unroll
                    for (int m = 0; m < NO; ++m)
                        if (NO <= 3 || m < p.nother) x *= u[b][m][vv].v[e];
                    contrib[vv * VEC + e] = x;
                }
            if (rws[b] >= 0) {
                if (rws[b] != cur_row) {
                    flush(false);
#pragma unroll
                    for (int c = 0; c < CPL; ++c) acc[c] = 0.0;
                    cur_row = rws[b];
                }
#pragma unroll
                for (int c = 0; c < CPL; ++c) acc[c] += contrib[c];
            }
        }
    }
    if (start < end) flush(true);
    __syncthreads();

    // ---- merge the 2*NG group carries of this CTA in slot order
    // run starts are handled by the group that owns the slot; a run is written to M unless it
    // contains the first or the last valid slot of the CTA (those go to the global carry list).
    int last_valid = -1;
    for (int k = 2 * NG - 1; k >= 0; --k)
        if (crow[k] >= 0) { last_valid = k; break; }
    if (last_valid < 0) {
        if (tid == 0) { p.carry_row[2 * blockIdx.x] = -1; p.carry_row[2 * blockIdx.x + 1] = -1; }
        return;
    }
    if (tid == 0) p.carry_row[2 * blockIdx.x + 1] = -1;  // overwritten below if a distinct tail run exists
    __syncthreads();
    for (int sl = 0; sl < 2; ++sl) {
        const int k = 2 * gi + sl;
        const int row = crow[k];
        if (row < 0) continue;
        int prev = k - 1;
        while (prev >= 0 && crow[prev] < 0) --prev;
        if (prev >= 0 && crow[prev] == row) continue;  // not a run start
        double sum[CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) sum[c] = 0.0;
        int j = k, last = k;
        while (j < 2 * NG && (crow[j] == row || crow[j] < 0)) {
            if (crow[j] == row) {
#pragma unroll
                for (int c = 0; c < CPL; ++c) sum[c] += cval[j * RW + colof(c)];
                last = j;
            }
            ++j;
        }
        const bool is_head = (prev < 0);
        const bool is_tail = (last == last_valid);
        if (is_head || is_tail) {
            const long long slot = 2LL * blockIdx.x + (is_head ? 0 : 1);
            if (gl == 0) p.carry_row[slot] = row;
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                const int col = colof(c);
                if (col < R) p.carry_val[slot * R + col] = sum[c];
            }
        } else {
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                const int col = colof(c);
                if (col < R) p.M[(long long)row * R + col] = sum[c];
            }
        }
    }
}

// Merge the per-CTA carries (2 per CTA, in storage order) and write the finished rows.
__global__ void coo_fixup_kernel(const int *__restrict__ carry_row, const double *__restrict__ carry_val,
                                 long long nslots, int R, double *M) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long k = t / R;
    const int col = (int)(t % R);
    if (k >= nslots) return;
    const int row = carry_row[k];
    if (row < 0) return;
    long long prev = k - 1;
    while (prev >= 0 && carry_row[prev] < 0) --prev;
    if (prev >= 0 && carry_row[prev] == row) return;
    double sum = 0.0;
    for (long long j = k; j < nslots; ++j) {
        const int rj = carry_row[j];
        if (rj < 0) continue;
        if (rj != row) break;
        sum += carry_val[j * R + col];
    }
    M[(long long)row * R + col] = sum;
}

// Independent check kernel: one thread per (non-zero, column), FP64 atomics on M.
struct AtomicParams {
    const int *idx[SKT_MAX_NDIM];
    const double *U[SKT_MAX_NDIM];
    const double *vals;
    long long nnz;
    int ndim, R, mode;
    int vmode, vdim;    // subscripts of mode `vmode` are stored as virtual rows panel * vdim + i (-1: none)
    double *M;
};
__global__ void mttkrp_coo_atomic_kernel(const __grid_constant__ AtomicParams p) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long z = t / p.R;
    const int r = (int)(t % p.R);
    if (z >= p.nnz) return;
    double v = p.vals[z];
    for (int m = 0; m < p.ndim; ++m)
        if (m != p.mode) v *= p.U[m][(long long)(m == p.vmode ? p.idx[m][z] % p.vdim : p.idx[m][z]) * p.R + r];
    const int orow = (p.mode == p.vmode) ? p.idx[p.mode][z] % p.vdim : p.idx[p.mode][z];
    atomicAdd(&p.M[(long long)orow * p.R + r], v);
}

// <P, X> partial sums: each thread owns (non-zero, column) pairs in a fixed stride pattern.
struct InnerParams {
    const int *idx[SKT_MAX_NDIM];
    const double *U[SKT_MAX_NDIM];
    const double *vals;
    const double *lambda;
    long long nnz;
    int ndim, R;
    int vmode, vdim;
};
__global__ void __launch_bounds__(CT) coo_innerprod_kernel(const __grid_constant__ InnerParams p, double *out,
                                                           double *partials, unsigned int *counter) {
    __shared__ double red[CT / 32];
    const long long total = p.nnz * p.R;
    double s = 0.0;
    for (long long t = blockIdx.x * (long long)CT + threadIdx.x; t < total; t += (long long)gridDim.x * CT) {
        const long long z = t / p.R;
        const int r = (int)(t % p.R);
        double v = p.vals[z] * p.lambda[r];
        for (int m = 0; m < p.ndim; ++m) v *= p.U[m][(long long)(m == p.vmode ? p.idx[m][z] % p.vdim : p.idx[m][z]) * p.R + r];
        s += v;
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double b = 0.0;
        for (int w = 0; w < CT / 32; ++w) b += red[w];
        partials[blockIdx.x] = b;
    }
    if (last_block_ticket(counter, gridDim.x)) {
        if (threadIdx.x == 0) {
            double tot = 0.0;
            for (unsigned k = 0; k < gridDim.x; ++k) tot += partials[k];
            out[0] = tot;
        }
    }
}

__global__ void coo_convert_kernel(const long long *__restrict__ src, long long n, long long dim, int *__restrict__ dst,
                                   unsigned int *bad) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long v = src[i];
        if (v < 0 || v >= dim) atomicAdd(bad, 1u);
        dst[i] = (int)v;
    }
}
__global__ void iota_kernel(unsigned int *p, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        p[i] = (unsigned int)i;
}
template <typename T>
__global__ void gather_kernel(const T *__restrict__ src, const unsigned int *__restrict__ perm, long long n,
                              T *__restrict__ dst) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = src[perm[i]];
}

// sort key of the panelled copy: virtual row = (subscript of the panel mode / panel_rows) * dim + output subscript
__global__ void panel_key_kernel(const int *__restrict__ out_idx, const int *__restrict__ pan_idx, long long n, int dim,
                                 int panel_rows, int *__restrict__ key) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        key[i] = (pan_idx[i] / panel_rows) * dim + out_idx[i];
}
// M = sum over the panels' slabs, in panel order (deterministic)
__global__ void panel_sum_kernel(const double *__restrict__ slabs, long long per, int npanel, double *__restrict__ M) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per; i += (long long)gridDim.x * blockDim.x) {
        double s = slabs[i];
        for (int q = 1; q < npanel; ++q) s += slabs[(long long)q * per + i];
        M[i] = s;
    }
}

int stream_blocks(long long n) { return (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, 8LL * sm_count())); }

// lanes per group / vectors per lane / vector width for rank R (16-byte vectors need aligned factor rows)
int group_cfg(int R, bool aligned, int &gs, int &vpl, int &vec) {
    SKT_REQUIRE(R >= 1 && R <= 128, SKT_EUNSUPPORTED, "sparse MTTKRP supports rank <= 128, got %d", R);
    vec = (aligned && R % 2 == 0) ? 2 : 1;
    const int nvec = (R + vec - 1) / vec;        // vectors per factor row
    vpl = (nvec + 31) / 32;
    const int per = (nvec + vpl - 1) / vpl;
    gs = 8;                                      // >= B so that one batch's subscripts are loaded by one group pass
    while (gs < per) gs *= 2;
    SKT_REQUIRE(vpl <= 4, SKT_EUNSUPPORTED, "sparse MTTKRP supports rank <= %d, got %d", 128 * vec, R);
    return SKT_OK;
}

long long coo_nctas(const skt_coo *t, int R, bool aligned) {
    int gs, vpl, vec;
    if (group_cfg(R, aligned, gs, vpl, vec) != SKT_OK) return 0;
    const long long per_cta = (long long)(CT / gs) * S;
    return std::max<long long>(1, (t->nnz + per_cta - 1) / per_cta);
}

template <int GS, int VPL, int VEC>
int launch_coo_no(int nother, const CooParams &p, long long nctas, cudaStream_t st) {
    constexpr int NG = CT / GS;
    const size_t smem = (size_t)((2 * NG * sizeof(int) + 7) / 8 * 8) + (size_t)2 * NG * GS * VPL * VEC * sizeof(double);
    const unsigned grid = (unsigned)nctas;
    if (nother <= 1) mttkrp_coo_kernel<GS, VPL, VEC, 1><<<grid, CT, smem, st>>>(p);
    else if (nother == 2) mttkrp_coo_kernel<GS, VPL, VEC, 2><<<grid, CT, smem, st>>>(p);
    else if (nother == 3) mttkrp_coo_kernel<GS, VPL, VEC, 3><<<grid, CT, smem, st>>>(p);
    else mttkrp_coo_kernel<GS, VPL, VEC, 7><<<grid, CT, smem, st>>>(p);
    SKT_LAUNCH_CHECK();
    return SKT_OK;
}

// cp.async-staged variant (16-byte aligned factors, even rank <= 64, at most 3 gathered modes)
template <int GS, int NO, int SBB, bool CA>
int launch_coo_staged_ca(const CooParams &p, long long nctas, cudaStream_t st) {
    constexpr int NG = CT / GS;
    const size_t carry = (size_t)((2 * NG * sizeof(int) + 7) / 8 * 8) + (size_t)2 * NG * GS * 2 * sizeof(double);
    const size_t smem = carry + (CT / 32) * staged_warp_bytes<GS, NO, SBB>();
    SKT_CUDA(cudaFuncSetAttribute(mttkrp_coo_kernel<GS, 1, 2, NO, true, SBB, CA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mttkrp_coo_kernel<GS, 1, 2, NO, true, SBB, CA><<<(unsigned)nctas, CT, smem, st>>>(p);
    SKT_LAUNCH_CHECK();
    return SKT_OK;
}
template <int GS, int NO, int SBB>
int launch_coo_staged_one(const CooParams &p, long long nctas, cudaStream_t st) {
    // SKT_COO_CA=1: gather the factor rows through L1 (cp.async.ca) -- an experiment switch; the default bypasses L1
    const char *e = getenv("SKT_COO_CA");
    if (e && e[0] == '1') return launch_coo_staged_ca<GS, NO, SBB, true>(p, nctas, st);
    return launch_coo_staged_ca<GS, NO, SBB, false>(p, nctas, st);
}

template <int GS>
int launch_coo_staged(int nother, const CooParams &p, long long nctas, cudaStream_t st) {
    static const char *e = getenv("SKT_COO_SBB");
    const bool small = !(e && atoi(e) == 8);   // 4 non-zeros per group and batch: 16 warps per SM (measured 6-10 % faster than 8)
    if (nother <= 1) return small ? launch_coo_staged_one<GS, 1, 4>(p, nctas, st) : launch_coo_staged_one<GS, 1, 8>(p, nctas, st);
    if (nother == 2) return small ? launch_coo_staged_one<GS, 2, 4>(p, nctas, st) : launch_coo_staged_one<GS, 2, 8>(p, nctas, st);
    return small ? launch_coo_staged_one<GS, 3, 4>(p, nctas, st) : launch_coo_staged_one<GS, 3, 8>(p, nctas, st);
}

bool staged_ok(int gs, int vpl, int vec, int nother) {
    const char *off = getenv("SKT_COO_NO_STAGED");
    if (off && off[0] == '1') return false;
    if (vec != 2 || vpl != 1 || nother > 3) return false;
    const size_t ring = (size_t)(CT / 32) * SD * ((32 / gs) * 8 * nother * gs * 16 + (32 / gs) * (8 * 16 + 16));
    return ring + 16384 <= smem_optin_bytes();
}

template <int VEC>
int launch_coo_vec(int gs, int vpl, int nother, const CooParams &p, long long nctas, cudaStream_t st) {
    if (vpl == 1) {
        switch (gs) {
            case 8: return launch_coo_no<8, 1, VEC>(nother, p, nctas, st);
            case 16: return launch_coo_no<16, 1, VEC>(nother, p, nctas, st);
            default: return launch_coo_no<32, 1, VEC>(nother, p, nctas, st);
        }
    }
    if (vpl == 2) return launch_coo_no<32, 2, VEC>(nother, p, nctas, st);
    if (vpl == 3) return launch_coo_no<32, 3, VEC>(nother, p, nctas, st);
    return launch_coo_no<32, 4, VEC>(nother, p, nctas, st);
}

int launch_coo(int gs, int vpl, int vec, int nother, const CooParams &p, long long nctas, cudaStream_t st) {
    if (staged_ok(gs, vpl, vec, nother)) {
        switch (gs) {
            case 8: return launch_coo_staged<8>(nother, p, nctas, st);
            case 16: return launch_coo_staged<16>(nother, p, nctas, st);
            default: return launch_coo_staged<32>(nother, p, nctas, st);
        }
    }
    return (vec == 2) ? launch_coo_vec<2>(gs, vpl, nother, p, nctas, st) : launch_coo_vec<1>(gs, vpl, nother, p, nctas, st);
}

int first_built(const skt_coo *t) {
    for (int n = 0; n < t->ndim; ++n)
        if (t->built_mask & (1u << n)) return n;
    return -1;
}

}  // namespace
}  // namespace skt

using namespace skt;

extern "C" int skt_coo_destroy(skt_coo *t) {
    if (!t) return SKT_OK;
    for (int n = 0; n < SKT_MAX_NDIM; ++n) {
        for (int m = 0; m < SKT_MAX_NDIM; ++m)
            if (t->idx[n][m]) cudaFree(t->idx[n][m]);
        if (t->vals[n]) cudaFree(t->vals[n]);
    }
    delete t;
    return SKT_OK;
}

extern "C" int64_t skt_coo_nnz(const skt_coo *t) { return t ? t->nnz : 0; }
extern "C" size_t skt_coo_device_bytes(const skt_coo *t) { return t ? t->bytes : 0; }

extern "C" int skt_coo_build(const int64_t *const *subs_d, const double *vals_d, int64_t nnz, const int64_t *shape,
                             int ndim, unsigned modes_mask, skt_stream stream, skt_coo **out) {
    SKT_REQUIRE(subs_d && shape && out && (vals_d || nnz == 0), SKT_EINVAL, "NULL argument");
    SKT_REQUIRE(ndim >= 2 && ndim <= SKT_MAX_NDIM, SKT_EINVAL, "ndim must be in [2, %d], got %d", SKT_MAX_NDIM, ndim);
    SKT_REQUIRE(nnz >= 0 && nnz < (1LL << 31) - 1, SKT_ESHAPE, "nnz %lld outside [0, 2^31)", (long long)nnz);
    for (int m = 0; m < ndim; ++m)
        SKT_REQUIRE((shape[m] >= 1 || (shape[m] == 0 && nnz == 0)) && shape[m] < (1LL << 31), SKT_ESHAPE,
                    "shape[%d] = %lld not in [1, 2^31)", m, (long long)shape[m]);
    cudaStream_t st = (cudaStream_t)stream;
    if (modes_mask == 0) modes_mask = (1u << ndim) - 1u;

    skt_coo *t = new skt_coo();
    memset(t, 0, sizeof(*t));
    t->ndim = ndim;
    t->nnz = nnz;
    for (int m = 0; m < ndim; ++m) t->shape[m] = shape[m];
    *out = nullptr;

    const size_t n = (size_t)std::max<int64_t>(nnz, 1);
    int *base[SKT_MAX_NDIM] = {nullptr};
    int *keys_out = nullptr;
    unsigned int *perm_in = nullptr, *perm_out = nullptr, *bad = nullptr;
    void *tmp = nullptr;
    int *pkey = nullptr;
    size_t tmp_bytes = 0;
    int rc = SKT_OK;
    auto cleanup = [&]() {
        for (int m = 0; m < ndim; ++m) cudaFree(base[m]);
        cudaFree(keys_out); cudaFree(perm_in); cudaFree(perm_out); cudaFree(bad); cudaFree(tmp); cudaFree(pkey);
    };
#define BUILD_TRY(call)                                                                \
    do {                                                                               \
        cudaError_t e__ = (call);                                                      \
        if (e__ != cudaSuccess) {                                                      \
            rc = cuda_fail(e__, #call, __FILE__, __LINE__);                            \
            cleanup(); skt_coo_destroy(t);                                             \
            return rc;                                                                 \
        }                                                                              \
    } while (0)
    BUILD_TRY(cudaMalloc(&bad, sizeof(unsigned int)));
    BUILD_TRY(cudaMemsetAsync(bad, 0, sizeof(unsigned int), st));
    for (int m = 0; m < ndim; ++m) {
        BUILD_TRY(cudaMalloc(&base[m], n * sizeof(int)));
        if (nnz > 0) coo_convert_kernel<<<stream_blocks(nnz), 256, 0, st>>>((const long long *)subs_d[m], nnz, shape[m], base[m], bad);
    }
    unsigned int nbad = 0;
    BUILD_TRY(cudaMemcpyAsync(&nbad, bad, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    BUILD_TRY(cudaStreamSynchronize(st));
    if (nbad) {
        cleanup(); skt_coo_destroy(t);
        set_error("%u subscripts fall outside the tensor shape", nbad);
        return SKT_EINVAL;
    }
    BUILD_TRY(cudaMalloc(&keys_out, n * sizeof(int)));
    BUILD_TRY(cudaMalloc(&perm_in, n * sizeof(unsigned int)));
    BUILD_TRY(cudaMalloc(&perm_out, n * sizeof(unsigned int)));
    BUILD_TRY(cudaMalloc(&pkey, n * sizeof(int)));
    tmp_bytes = rsort::scratch_bytes((long long)n);
    BUILD_TRY(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 256));

    // Index-panel blocking.  A gathered factor much larger than L2 (config 3: the 10 M x 16 factor of mode 0, 1.28 GB) is
    // re-fetched from DRAM for almost every non-zero when the copy is sorted by the output subscript alone (ncu, round 1:
    // 37 GB of DRAM reads for 5.4 GB of algorithmic bytes on mode 2).  Sorting by (panel of that mode, output subscript)
    // walks the big factor one L2-sized panel at a time; every panel owns a slab of M ("virtual rows"
    // panel * I_n + i_n -- the segmented-reduction kernel is unchanged) and the slabs are summed in panel order.  The
    // number of panels is capped so that the slab traffic stays small against the non-zero stream.
    long long want_rows = 1LL << 18;          // 32 MB of rank-16 FP64 factor rows per panel (measured: 2^18 5.56 ms, 2^19 5.74, 2^20 6.13, none 6.99 on mode 2 of config 3)
    long long cap_div = 32;                   // panels * I_n <= nnz / cap_div
    if (const char *e = getenv("SKT_COO_PANEL_ROWS")) want_rows = std::max<long long>(1, atoll(e));
    if (const char *e = getenv("SKT_COO_PANEL_CAP")) cap_div = std::max<long long>(0, atoll(e));
    const char *pan_off = getenv("SKT_COO_PANELS");
    const bool panels_on = !(pan_off && pan_off[0] == '0');
    for (int nmode = 0; nmode < ndim; ++nmode) {
        if (!(modes_mask & (1u << nmode))) continue;
        int g = -1;
        for (int m = 0; m < ndim; ++m)
            if (m != nmode && (g < 0 || shape[m] > shape[g])) g = m;
        long long P = panels_on && nnz > 0 ? (shape[g] + want_rows - 1) / want_rows : 1;
        if (cap_div > 0) P = std::min<long long>(P, (nnz / cap_div) / std::max<long long>(shape[nmode], 1));
        P = std::min<long long>(P, ((1LL << 31) - 2) / std::max<long long>(shape[nmode], 1));
        if (P < 2) P = 1;
        const long long prow = (shape[g] + P - 1) / P;
        t->npanel[nmode] = (int)P;
        t->panel_mode[nmode] = g;
        t->panel_rows[nmode] = prow;
        int bits = 1;
        while ((1LL << bits) < shape[nmode] * P) ++bits;
        const int *sorted_keys = keys_out;
        const unsigned int *sorted_perm = perm_out;
        if (nnz > 0) {
            // sort key -> pkey (the sort ping-pongs between (pkey, perm_in) and (keys_out, perm_out))
            if (P > 1) {
                panel_key_kernel<<<stream_blocks(nnz), 256, 0, st>>>(base[nmode], base[g], nnz, (int)shape[nmode], (int)prow, pkey);
            } else {
                BUILD_TRY(cudaMemcpyAsync(pkey, base[nmode], (size_t)nnz * sizeof(int), cudaMemcpyDeviceToDevice, st));
            }
            iota_kernel<<<stream_blocks(nnz), 256, 0, st>>>(perm_in, nnz);
            int where = 0;
            rc = rsort::sort_pairs<unsigned int>((unsigned int *)pkey, perm_in, (unsigned int *)keys_out, perm_out, nnz, bits,
                                                 tmp, st, &where);
            if (rc != SKT_OK) { cleanup(); skt_coo_destroy(t); return rc; }
            if (where == 0) { sorted_keys = pkey; sorted_perm = perm_in; }
        }
        for (int m = 0; m < ndim; ++m) {
            BUILD_TRY(cudaMalloc(&t->idx[nmode][m], n * sizeof(int)));
            t->bytes += n * sizeof(int);
            if (nnz > 0) {
                if (m == nmode)               // the sorted keys ARE the (virtual) rows
                    BUILD_TRY(cudaMemcpyAsync(t->idx[nmode][m], sorted_keys, (size_t)nnz * sizeof(int), cudaMemcpyDeviceToDevice, st));
                else
                    gather_kernel<int><<<stream_blocks(nnz), 256, 0, st>>>(base[m], sorted_perm, nnz, t->idx[nmode][m]);
            }
        }
        BUILD_TRY(cudaMalloc(&t->vals[nmode], n * sizeof(double)));
        t->bytes += n * sizeof(double);
        if (nnz > 0) gather_kernel<double><<<stream_blocks(nnz), 256, 0, st>>>(vals_d, sorted_perm, nnz, t->vals[nmode]);
        t->built_mask |= (1u << nmode);
    }
    BUILD_TRY(cudaGetLastError());
    BUILD_TRY(cudaStreamSynchronize(st));
#undef BUILD_TRY
    cleanup();
    *out = t;
    return SKT_OK;
}

namespace {
struct PhaseTimer {
    bool on; double t0;
    static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
    PhaseTimer() { const char *e = getenv("SKT_COO_TRACE"); on = e && e[0] == '1'; t0 = now(); }
    void lap(const char *what) { if (on) { cudaDeviceSynchronize(); const double t = now(); fprintf(stderr, "[coo build] %-28s %8.1f ms\n", what, (t - t0) * 1e3); t0 = t; } }
};
}  // namespace

extern "C" int skt_coo_build_host(const int64_t *const *subs_h, const double *vals_h, int64_t nnz,
                                  const int64_t *shape, int ndim, unsigned modes_mask, skt_coo **out) {
    SKT_REQUIRE(subs_h && shape && out && (vals_h || nnz == 0), SKT_EINVAL, "NULL argument");
    SKT_REQUIRE(ndim >= 2 && ndim <= SKT_MAX_NDIM, SKT_EINVAL, "ndim must be in [2, %d], got %d", SKT_MAX_NDIM, ndim);
    SKT_REQUIRE(nnz >= 0, SKT_ESHAPE, "negative nnz");
    const size_t n = (size_t)std::max<int64_t>(nnz, 1);
    long long *subs_d[SKT_MAX_NDIM] = {nullptr};
    double *vals_d = nullptr;
    PhaseTimer pt_start;
    auto cleanup = [&]() {
        for (int m = 0; m < ndim; ++m) cudaFree(subs_d[m]);
        cudaFree(vals_d);
    };
    // the COO arrays come from pageable NumPy memory: multi-threaded pinned staging (skt_upload, ~4x a plain cudaMemcpy)
    for (int m = 0; m < ndim; ++m) {
        cudaError_t e = cudaMalloc(&subs_d[m], n * sizeof(long long));
        if (e != cudaSuccess) { cleanup(); return cuda_fail(e, "allocate subscripts", __FILE__, __LINE__); }
        if (nnz > 0) {
            const int rc = skt_upload(subs_d[m], subs_h[m], (size_t)nnz * sizeof(long long), nullptr);
            if (rc != SKT_OK) { cleanup(); return rc; }
        }
    }
    cudaError_t e = cudaMalloc(&vals_d, n * sizeof(double));
    if (e != cudaSuccess) { cleanup(); return cuda_fail(e, "allocate values", __FILE__, __LINE__); }
    if (nnz > 0) {
        const int rc = skt_upload(vals_d, vals_h, (size_t)nnz * sizeof(double), nullptr);
        if (rc != SKT_OK) { cleanup(); return rc; }
    }
    e = cudaStreamSynchronize(nullptr);
    if (e != cudaSuccess) { cleanup(); return cuda_fail(e, "upload", __FILE__, __LINE__); }
    pt_start.lap("malloc + upload (host)");
    PhaseTimer pt_after_upload;
    const int rc = skt_coo_build((const int64_t *const *)subs_d, vals_d, nnz, shape, ndim, modes_mask, nullptr, out);
    pt_after_upload.lap("sort + gather (device)");
    cleanup();
    pt_after_upload.lap("free staging");
    return rc;
}

extern "C" size_t skt_mttkrp_coo_workspace(const skt_coo *t, int rank, int mode) {
    if (!t || rank < 1) return 0;
    const long long nctas = coo_nctas(t, rank, false);   // the scalar layout needs the most CTAs: covers both
    size_t slabs = 0;                                     // panelled copy: one slab of M per panel
    if (mode >= 0 && mode < t->ndim && t->npanel[mode] > 1)
        slabs = align_up((size_t)t->npanel[mode] * t->shape[mode] * rank * sizeof(double), 256);
    return align_up((size_t)2 * nctas * sizeof(int), 256) + align_up((size_t)2 * nctas * rank * sizeof(double), 256) + slabs;
}

extern "C" int skt_mttkrp_coo(const skt_coo *t, const double *const *U_d, int rank, int mode, double *M_d, void *ws_d,
                              size_t ws_bytes, skt_stream stream) {
    SKT_REQUIRE(t && U_d, SKT_EINVAL, "NULL argument");
    SKT_REQUIRE(mode >= 0 && mode < t->ndim, SKT_EINVAL, "mode %d out of range", mode);
    SKT_REQUIRE(rank >= 1, SKT_EINVAL, "rank must be >= 1");
    if (t->nnz == 0) {   // empty tensor (a rank of a sharded run may own no non-zeros): all zeros
        if (t->shape[mode] > 0) {
            SKT_REQUIRE(M_d, SKT_EINVAL, "NULL output");
            SKT_CUDA(cudaMemsetAsync(M_d, 0, (size_t)t->shape[mode] * rank * sizeof(double), (cudaStream_t)stream));
        }
        return SKT_OK;
    }
    SKT_REQUIRE(M_d, SKT_EINVAL, "NULL output");
    SKT_REQUIRE(t->built_mask & (1u << mode), SKT_EINVAL, "the sorted copy for mode %d was not built", mode);
    bool aligned = true;                     // 16-byte factor-row loads need 16-byte aligned factors
    for (int m = 0; m < t->ndim; ++m)
        if (m != mode && U_d[m] && (((uintptr_t)U_d[m]) & 15) != 0) aligned = false;
    int gs, vpl, vec;
    int rc = group_cfg(rank, aligned, gs, vpl, vec);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t need = skt_mttkrp_coo_workspace(t, rank, mode);
    SKT_REQUIRE(ws_d && ws_bytes >= need, SKT_EWORKSPACE, "workspace too small: need %zu bytes, got %zu", need, ws_bytes);
    const long long nctas = coo_nctas(t, rank, aligned);
    const long long nctas_ws = coo_nctas(t, rank, false);
    const int npanel = t->npanel[mode];
    const long long per = t->shape[mode] * (long long)rank;
    CooParams p;
    p.carry_row = (int *)ws_d;
    p.carry_val = (double *)((char *)ws_d + align_up((size_t)2 * nctas_ws * sizeof(int), 256));
    // panelled copy: the kernel writes virtual rows (panel * I_n + i_n) into per-panel slabs, summed below
    double *target = npanel > 1 ? (double *)((char *)p.carry_val + align_up((size_t)2 * nctas_ws * rank * sizeof(double), 256)) : M_d;
    SKT_CUDA(cudaMemsetAsync(target, 0, (size_t)npanel * per * sizeof(double), st));
    p.rows = t->idx[mode][mode];
    p.vals = t->vals[mode];
    p.nnz = t->nnz;
    p.R = rank;
    p.M = target;
    int k = 0;
    for (int m = 0; m < t->ndim; ++m) {
        if (m == mode) continue;
        SKT_REQUIRE(U_d[m] != nullptr, SKT_EINVAL, "U[%d] is NULL (only U[mode] may be)", m);
        p.oidx[k] = t->idx[mode][m];
        p.U[k] = U_d[m];
        ++k;
    }
    p.nother = k;
    for (; k < SKT_MAX_NDIM; ++k) { p.oidx[k] = p.oidx[0]; p.U[k] = p.U[0]; }
    rc = launch_coo(gs, vpl, vec, p.nother, p, nctas, st);
    if (rc) return rc;
    const long long nslots = 2 * nctas;
    const long long threads = nslots * rank;
    coo_fixup_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(p.carry_row, p.carry_val, nslots, rank, target);
    SKT_LAUNCH_CHECK();
    if (npanel > 1) {
        panel_sum_kernel<<<stream_blocks(per), 256, 0, st>>>(target, per, npanel, M_d);
        SKT_LAUNCH_CHECK();
    }
    return SKT_OK;
}

extern "C" int skt_mttkrp_coo_atomic(const skt_coo *t, const double *const *U_d, int rank, int mode, double *M_d,
                                     skt_stream stream) {
    SKT_REQUIRE(t && U_d && M_d, SKT_EINVAL, "NULL argument");
    SKT_REQUIRE(mode >= 0 && mode < t->ndim && rank >= 1, SKT_EINVAL, "bad mode/rank");
    const int src = first_built(t);
    SKT_REQUIRE(src >= 0, SKT_EINVAL, "no sorted copy was built");
    cudaStream_t st = (cudaStream_t)stream;
    SKT_CUDA(cudaMemsetAsync(M_d, 0, (size_t)t->shape[mode] * rank * sizeof(double), st));
    if (t->nnz == 0) return SKT_OK;
    AtomicParams p;
    for (int m = 0; m < t->ndim; ++m) {
        SKT_REQUIRE(m == mode || U_d[m] != nullptr, SKT_EINVAL, "U[%d] is NULL", m);
        p.idx[m] = t->idx[src][m];
        p.U[m] = U_d[m];
    }
    p.vals = t->vals[src];
    p.nnz = t->nnz; p.ndim = t->ndim; p.R = rank; p.mode = mode; p.M = M_d;
    p.vmode = t->npanel[src] > 1 ? src : -1; p.vdim = (int)t->shape[src];
    const long long threads = t->nnz * rank;
    mttkrp_coo_atomic_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(p);
    SKT_LAUNCH_CHECK();
    return SKT_OK;
}

extern "C" size_t skt_coo_innerprod_workspace(const skt_coo *t, int rank) {
    (void)t; (void)rank;
    return 256 + align_up((size_t)4 * sm_count() * sizeof(double), 256);
}

extern "C" int skt_coo_innerprod(const skt_coo *t, const double *const *U_d, const double *lambda_d, int rank,
                                 double *out_d, void *ws_d, size_t ws_bytes, skt_stream stream) {
    SKT_REQUIRE(t && U_d && lambda_d && out_d, SKT_EINVAL, "NULL argument");
    SKT_REQUIRE(rank >= 1, SKT_EINVAL, "rank must be >= 1");
    const int src = first_built(t);
    SKT_REQUIRE(src >= 0, SKT_EINVAL, "no sorted copy was built");
    const size_t need = skt_coo_innerprod_workspace(t, rank);
    SKT_REQUIRE(ws_d && ws_bytes >= need, SKT_EWORKSPACE, "workspace too small: need %zu bytes, got %zu", need, ws_bytes);
    cudaStream_t st = (cudaStream_t)stream;
    InnerParams p;
    for (int m = 0; m < t->ndim; ++m) {
        SKT_REQUIRE(U_d[m] != nullptr, SKT_EINVAL, "U[%d] is NULL", m);
        p.idx[m] = t->idx[src][m];
        p.U[m] = U_d[m];
    }
    p.vals = t->vals[src]; p.lambda = lambda_d; p.nnz = t->nnz; p.ndim = t->ndim; p.R = rank;
    p.vmode = t->npanel[src] > 1 ? src : -1; p.vdim = (int)t->shape[src];
    unsigned int *counter = (unsigned int *)ws_d;
    double *partials = (double *)((char *)ws_d + 256);
    SKT_CUDA(cudaMemsetAsync(counter, 0, 256, st));
    const long long total = std::max<long long>(1, t->nnz * rank);
    const int nb = (int)std::max<long long>(1, std::min<long long>((total + CT - 1) / CT, 4LL * sm_count()));
    coo_innerprod_kernel<<<nb, CT, 0, st>>>(p, out_d, partials, counter);
    SKT_LAUNCH_CHECK();
    return SKT_OK;
}

extern "C" int skt_coo_sumsq(const skt_coo *t, double *out_d, void *ws_d, size_t ws_bytes, skt_stream stream) {
    SKT_REQUIRE(t && out_d, SKT_EINVAL, "NULL argument");
    const int src = first_built(t);
    SKT_REQUIRE(src >= 0, SKT_EINVAL, "no sorted copy was built");
    return skt_sumsq(t->vals[src], t->nnz, out_d, ws_d, ws_bytes, stream);
}

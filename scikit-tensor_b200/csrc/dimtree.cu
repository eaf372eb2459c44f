// Dimension-tree reuse for the dense CP-ALS sweep.
//
// The reference runs one full MTTKRP per mode (sktensor/cp.py:142-143 -> dtensor.py:163-167), i.e. N
// passes over X per ALS iteration.  The factors of the modes right of a split point s do not change
// while the modes left of it are updated, so one partial contraction
//
//     T_A[i_0..i_{s-1}, r] = sum_{i_s..i_{N-1}} X[i] * prod_{m >= s} U_m[i_m, r]
//
// serves every MTTKRP of the left group, and T_B (the mirror image) every MTTKRP of the right group:
// two passes over X per iteration for any N.  T_A / T_B are produced by skt_mttkrp_dense itself on a
// reshaped view of X ((prod_{m<s} I_m) x I_s x ... mode 0, resp. I_0 x ... x (prod_{m>=s} I_m) last
// mode) -- no new tensor kernel.  This file holds the small second stage,
//
//     M_n[i_n, r] = sum_{other i of the group} T[i_group, r] * prod_{m in group, m != n} U_m[i_m, r],
//
// a streaming, HBM/L2-bound reduction over the (small) partial tensor T, with a fixed summation order
// (per-split partial slabs added in split order: no atomics, bitwise reproducible).
#include <algorithm>

#include "common.cuh"

namespace skt {
namespace {

struct KrMode {
    const double *W;  // factor of a group mode other than the output mode
    int dim, div, space;  // index = (space_index / div) % dim ; space 0 = left of the output mode, 1 = right
};

struct KrParams {
    const double *T;
    double *out;              // M (nsplit == 1) or partial slabs [nsplit][In][R]
    long long In, L, Tt, J;   // J = L * Tt terms per output
    long long chunk;
    int R, nsplit, nw;
    KrMode w[SKT_MAX_NDIM];
};

// One thread per (output row, column, split); consecutive threads own consecutive columns, so a warp
// reads contiguous 8*R-byte rows of T.  NW1: exactly one weight mode (the common two-mode group).
template <bool NW1>
__global__ void __launch_bounds__(256) kr_reduce_kernel(const __grid_constant__ KrParams p) {
    const long long e = blockIdx.x * 256LL + threadIdx.x;
    if (e >= p.In * p.R) return;
    const long long i = e / p.R;
    const int r = (int)(e - i * p.R);
    const int s = blockIdx.y;
    const long long j0 = s * p.chunk, j1 = min(j0 + p.chunk, p.J);
    double sum = 0.0;
    if (NW1) {
        const KrMode w = p.w[0];
        // left weight: j = l (Tt == 1), T row (l*In + i); right weight: j = t (L == 1), T row (i*Tt + t)
        const long long stride = w.space ? (long long)p.R : p.In * p.R;
        const double *tp = p.T + (w.space ? (i * p.Tt + j0) * p.R : (j0 * p.In + i) * p.R) + r;
        const double *wp = w.W + j0 * p.R + r;
        long long j = j0;
        // 4 independent loads of T per thread in flight (8 measured slower: 22.6 vs 15.5 us at 512 x 512 x 32)
        for (; j + 4 <= j1; j += 4) {
            const double t0 = __ldg(tp), t1 = __ldg(tp + stride), t2 = __ldg(tp + 2 * stride), t3 = __ldg(tp + 3 * stride);
            const double w0 = __ldg(wp), w1 = __ldg(wp + p.R), w2 = __ldg(wp + 2 * p.R), w3 = __ldg(wp + 3 * p.R);
            sum += t0 * w0;
            sum += t1 * w1;
            sum += t2 * w2;
            sum += t3 * w3;
            tp += 4 * stride;
            wp += 4 * p.R;
        }
        for (; j < j1; ++j) {
            sum += __ldg(tp) * __ldg(wp);
            tp += stride;
            wp += p.R;
        }
    } else {
        for (long long j = j0; j < j1; ++j) {
            const long long l = j / p.Tt, t = j - l * p.Tt;
            double w = 1.0;
            for (int m = 0; m < p.nw; ++m) {
                const long long sp = p.w[m].space ? t : l;
                const long long idx = (sp / p.w[m].div) % p.w[m].dim;
                w *= __ldg(p.w[m].W + idx * p.R + r);
            }
            sum += __ldg(p.T + ((l * p.In + i) * p.Tt + t) * p.R + r) * w;
        }
    }
    p.out[(long long)s * p.In * p.R + e] = sum;
}

__global__ void kr_sum_splits_kernel(const double *__restrict__ P, int nsplit, long long n, double *__restrict__ M) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int k = 0; k < nsplit; ++k) s += P[(long long)k * n + e];
        M[e] = s;
    }
}

int kr_plan(const int64_t *dims, int p, int rank, int mode, long long &In, long long &L, long long &Tt, int &nsplit,
            long long &chunk) {
    SKT_REQUIRE(dims != nullptr, SKT_EINVAL, "dims is NULL");
    SKT_REQUIRE(p >= 2 && p <= SKT_MAX_NDIM, SKT_EINVAL, "group size must be in [2, %d], got %d", SKT_MAX_NDIM, p);
    SKT_REQUIRE(mode >= 0 && mode < p && rank >= 1, SKT_EINVAL, "bad mode %d / rank %d", mode, rank);
    L = 1; Tt = 1;
    for (int m = 0; m < p; ++m) {
        SKT_REQUIRE(dims[m] >= 1 && dims[m] < (1LL << 31), SKT_ESHAPE, "dims[%d] = %lld", m, (long long)dims[m]);
        if (m < mode) L *= dims[m];
        if (m > mode) Tt *= dims[m];
        SKT_REQUIRE(L < (1LL << 40) && Tt < (1LL << 40), SKT_ESHAPE, "partial tensor too large");
    }
    In = dims[mode];
    const long long J = L * Tt;
    // enough CTAs to fill the chip (>= 4 per SM) without making the chunks trivially short
    const long long ctas = (In * rank + 255) / 256;
    long long ns = std::max<long long>(1, (4LL * sm_count() + ctas - 1) / ctas);
    ns = std::min<long long>(ns, std::max<long long>(1, J / 16));
    ns = std::min<long long>(ns, 1024);
    chunk = (J + ns - 1) / ns;
    nsplit = (int)((J + chunk - 1) / chunk);
    if (nsplit < 1) nsplit = 1;
    return SKT_OK;
}


// ---------------------------------------------------------------------------------------------
// Dense reconstruction of a Kruskal tensor (ktensor.toarray, sktensor/ktensor.py:152-163):
//   out[i_0..i_{N-1}] = sum_r lambda_r prod_m U_m[i_m, r]        (C order)
// One CTA per index of all modes but the last: the R prefix products go to shared memory, the threads then sweep
// the last mode.  The factor of the last mode (I x R) stays in L1/L2; the output is written once, coalesced.
// ---------------------------------------------------------------------------------------------
struct FullParams {
    const double *U[SKT_MAX_NDIM];
    const double *lambda;
    long long shape[SKT_MAX_NDIM];
    double *out;
    long long nprefix;
    int ndim, R;
};

__global__ void __launch_bounds__(256) ktensor_full_kernel(const __grid_constant__ FullParams p) {
    extern __shared__ double pre[];
    const int last = p.ndim - 1;
    const long long K = p.shape[last];
    for (long long q = blockIdx.x; q < p.nprefix; q += gridDim.x) {
        __syncthreads();
        for (int r = threadIdx.x; r < p.R; r += 256) {
            double v = p.lambda ? p.lambda[r] : 1.0;
            long long rem = q;
            for (int m = last - 1; m >= 0; --m) {
                const long long im = rem % p.shape[m];
                rem /= p.shape[m];
                v *= __ldg(p.U[m] + im * p.R + r);
            }
            pre[r] = v;
        }
        __syncthreads();
        for (long long k = threadIdx.x; k < K; k += 256) {
            const double *row = p.U[last] + k * p.R;
            double sum = 0.0;
            for (int r = 0; r < p.R; ++r) sum += pre[r] * __ldg(row + r);
            p.out[q * K + k] = sum;
        }
    }
}

}  // namespace
}  // namespace skt

using namespace skt;

extern "C" size_t skt_kr_reduce_workspace(const int64_t *dims, int p, int rank, int mode) {
    if (dims && p >= 2 && p <= SKT_MAX_NDIM)
        for (int m = 0; m < p; ++m)
            if (dims[m] == 0) return 256;
    long long In, L, Tt, chunk;
    int nsplit;
    if (kr_plan(dims, p, rank, mode, In, L, Tt, nsplit, chunk) != SKT_OK) return 0;
    return align_up((size_t)nsplit * In * rank * sizeof(double) + 256, 256);   // (also covers skt_kr_reduce_slabs with one split)
}

static int kr_reduce_impl(const double *T_d, const int64_t *dims, int p, const double *const *W_d, int rank, int mode,
                          double *M_d, void *ws_d, size_t ws_bytes, skt_stream stream, int *nslabs_out);

extern "C" int skt_kr_reduce(const double *T_d, const int64_t *dims, int p, const double *const *W_d, int rank, int mode,
                             double *M_d, void *ws_d, size_t ws_bytes, skt_stream stream) {
    return kr_reduce_impl(T_d, dims, p, W_d, rank, mode, M_d, ws_d, ws_bytes, stream, nullptr);
}

// Same, but the per-split partial slabs are left un-summed in slabs_d ([*nslabs_out][dims[mode]][rank], at most
// skt_kr_reduce_workspace bytes): skt_cp_update_cluster_slabs adds them (in split order, the same arithmetic) while it
// stages M, which saves a launch on the critical path of the sweep.
extern "C" int skt_kr_reduce_slabs(const double *T_d, const int64_t *dims, int p, const double *const *W_d, int rank,
                                   int mode, double *slabs_d, size_t slabs_bytes, int *nslabs_out, skt_stream stream) {
    SKT_REQUIRE(nslabs_out != nullptr, SKT_EINVAL, "NULL argument");
    return kr_reduce_impl(T_d, dims, p, W_d, rank, mode, nullptr, slabs_d, slabs_bytes, stream, nslabs_out);
}

static int kr_reduce_impl(const double *T_d, const int64_t *dims, int p, const double *const *W_d, int rank, int mode,
                          double *M_d, void *ws_d, size_t ws_bytes, skt_stream stream, int *nslabs_out) {
    const bool keep = nslabs_out != nullptr;      // leave the partial slabs in ws_d
    if (dims && p >= 2 && p <= SKT_MAX_NDIM && mode >= 0 && mode < p && rank >= 1) {
        bool empty = false;
        for (int m = 0; m < p; ++m) empty = empty || dims[m] == 0;
        if (empty) {     // empty group (a rank without rows): all zeros
            double *dst = keep ? (double *)ws_d : M_d;
            if (dims[mode] > 0 && dst) SKT_CUDA(cudaMemsetAsync(dst, 0, (size_t)dims[mode] * rank * sizeof(double), (cudaStream_t)stream));
            if (keep) *nslabs_out = 1;
            return SKT_OK;
        }
    }
    SKT_REQUIRE(T_d && W_d && (M_d || keep), SKT_EINVAL, "NULL argument");
    long long In, L, Tt, chunk;
    int nsplit;
    int rc = kr_plan(dims, p, rank, mode, In, L, Tt, nsplit, chunk);
    if (rc != SKT_OK) return rc;
    const size_t need = skt_kr_reduce_workspace(dims, p, rank, mode);
    SKT_REQUIRE((nsplit == 1 && !keep) || (ws_d && ws_bytes >= (keep ? (size_t)nsplit * In * rank * sizeof(double) : need)),
                SKT_EWORKSPACE, "workspace too small: need %zu bytes, got %zu", need, ws_bytes);
    KrParams kp;
    kp.T = T_d;
    kp.In = In; kp.L = L; kp.Tt = Tt; kp.J = L * Tt;
    kp.chunk = chunk; kp.R = rank; kp.nsplit = nsplit;
    kp.out = (nsplit > 1 || keep) ? (double *)ws_d : M_d;
    kp.nw = 0;
    {
        long long div = 1;
        for (int m = mode - 1; m >= 0; --m) {
            SKT_REQUIRE(W_d[m] != nullptr, SKT_EINVAL, "W[%d] is NULL (only W[mode] may be)", m);
            kp.w[kp.nw].W = W_d[m]; kp.w[kp.nw].dim = (int)dims[m]; kp.w[kp.nw].div = (int)div; kp.w[kp.nw].space = 0;
            ++kp.nw;
            div *= dims[m];
            SKT_REQUIRE(div < (1LL << 31), SKT_ESHAPE, "group too large");
        }
        div = 1;
        for (int m = p - 1; m > mode; --m) {
            SKT_REQUIRE(W_d[m] != nullptr, SKT_EINVAL, "W[%d] is NULL (only W[mode] may be)", m);
            kp.w[kp.nw].W = W_d[m]; kp.w[kp.nw].dim = (int)dims[m]; kp.w[kp.nw].div = (int)div; kp.w[kp.nw].space = 1;
            ++kp.nw;
            div *= dims[m];
            SKT_REQUIRE(div < (1LL << 31), SKT_ESHAPE, "group too large");
        }
    }
    cudaStream_t st = (cudaStream_t)stream;
    const long long ctas = (In * rank + 255) / 256;
    SKT_REQUIRE(ctas < (1LL << 31), SKT_ESHAPE, "output too large");
    dim3 grid((unsigned)ctas, (unsigned)nsplit);
    if (kp.nw == 1) kr_reduce_kernel<true><<<grid, 256, 0, st>>>(kp);
    else kr_reduce_kernel<false><<<grid, 256, 0, st>>>(kp);
    SKT_LAUNCH_CHECK();
    if (keep) {
        *nslabs_out = nsplit;
        return SKT_OK;
    }
    if (nsplit > 1) {
        const long long n = In * rank;
        const int blocks = (int)std::min<long long>((n + 255) / 256, 4LL * sm_count());
        kr_sum_splits_kernel<<<blocks, 256, 0, st>>>((const double *)ws_d, nsplit, n, M_d);
        SKT_LAUNCH_CHECK();
    }
    return SKT_OK;
}

extern "C" int skt_ktensor_full(const double *const *U_d, const double *lambda_d, const int64_t *shape, int ndim, int rank,
                                double *out_d, skt_stream stream) {
    SKT_REQUIRE(U_d && shape && out_d, SKT_EINVAL, "NULL argument");
    SKT_REQUIRE(ndim >= 1 && ndim <= SKT_MAX_NDIM, SKT_EINVAL, "ndim must be in [1, %d], got %d", SKT_MAX_NDIM, ndim);
    SKT_REQUIRE(rank >= 1 && rank <= 4096, SKT_EUNSUPPORTED, "reconstruction supports rank <= 4096, got %d", rank);
    FullParams p;
    p.lambda = lambda_d; p.out = out_d; p.ndim = ndim; p.R = rank;
    long long npre = 1;
    for (int m = 0; m < ndim; ++m) {
        SKT_REQUIRE(shape[m] >= 0 && shape[m] < (1LL << 31), SKT_ESHAPE, "shape[%d] = %lld", m, (long long)shape[m]);
        if (shape[m] == 0) return SKT_OK;              // empty tensor
        SKT_REQUIRE(U_d[m] != nullptr, SKT_EINVAL, "U[%d] is NULL", m);
        p.U[m] = U_d[m];
        p.shape[m] = shape[m];
        if (m < ndim - 1) npre *= shape[m];
        SKT_REQUIRE(npre < (1LL << 46), SKT_ESHAPE, "tensor too large");
    }
    p.nprefix = npre;
    const unsigned grid = (unsigned)std::min<long long>(npre, 64LL * sm_count());
    ktensor_full_kernel<<<grid, 256, (size_t)rank * sizeof(double), (cudaStream_t)stream>>>(p);
    SKT_LAUNCH_CHECK();
    return SKT_OK;
}

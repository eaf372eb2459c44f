// CP-ALS normal equations for rank > 64 (the reference's cp.als has no rank limit, sktensor/cp.py:46,144-159).
//
// The rank <= 64 kernels (cpals.cu, cpals_fused.cu, pinv.cuh) keep the R x R system in one CTA's registers / shared
// memory.  Wider systems are rare and never the bottleneck of a sweep, so they are assembled from the large-matrix
// kernels the library already has, behind the same entry points (skt_gram, skt_hadamard_pinv, skt_solve_stats,
// skt_normalise_gram):
//   G = U^T U                          -> skt_unfold_gram on the (rows x R) view, mode 1 (FP64 tensor cores)
//   pinv(Y), Y = hadamard of the Grams -> skt_symeig_top with k = R (Y is symmetric: its singular values are |lambda|),
//                                         P = V diag(1/lambda_i, |lambda_i| > R eps max|lambda|) V^T  -- scipy.linalg.pinv's
//                                         cut-off (cp.py:147)
//   Unew = M P                         -> skt_ttm_dense on the (rows x R) view, mode 1 (P is symmetric)
//   column statistics, U / lambda, <P,X> partials: the streaming kernels below (ordered reductions).
#include <algorithm>
#include <cfloat>

#include "common.cuh"

namespace skt {
namespace {

constexpr int WT = 256;

__global__ void wide_hadamard_kernel(const double *__restrict__ G, int ndim, int R, int mode, double *Y, int *info) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= R * R) return;
    double y = 1.0;
    for (int m = 0; m < ndim; ++m)
        if (m != mode) y *= G[(size_t)m * R * R + e];
    Y[e] = y;
    if (!(fabs(y) <= DBL_MAX)) atomicExch(info, -1);
}

// P = V diag(f) V^T with f_i = 1 / w_i for |w_i| > R eps max|w|, else 0.  V is R x R row-major (column i = eigenvector
// of w[i]); one CTA per 16 x 16 tile of P.
__global__ void wide_pinv_from_eig_kernel(const double *__restrict__ V, const double *__restrict__ W, int R, double *P,
                                          int *info) {
    extern __shared__ double f[];
    __shared__ double smax;
    __shared__ int kept;
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        double m = 0.0;
        for (int i = 0; i < R; ++i) m = fmax(m, fabs(W[i]));
        smax = m;
        kept = 0;
    }
    __syncthreads();
    const double cut = (double)R * DBL_EPSILON * smax;
    const int t = threadIdx.y * blockDim.x + threadIdx.x;
    for (int i = t; i < R; i += blockDim.x * blockDim.y) {
        const double w = W[i];
        const bool keep = fabs(w) > cut;
        f[i] = keep ? 1.0 / w : 0.0;
        if (keep && blockIdx.x == 0 && blockIdx.y == 0) atomicAdd(&kept, 1);
    }
    __syncthreads();
    const int a = blockIdx.y * blockDim.y + threadIdx.y, b = blockIdx.x * blockDim.x + threadIdx.x;
    const bool bad = info[0] < 0;
    if (a < R && b < R) {
        double s = 0.0;
        for (int i = 0; i < R; ++i) s += V[(size_t)a * R + i] * f[i] * V[(size_t)b * R + i];
        P[(size_t)a * R + b] = bad ? nan("") : s;
    }
    if (blockIdx.x == 0 && blockIdx.y == 0 && t == 0 && !bad) info[0] = kept;
}

// per-column sum of squares (first iteration) or signed maximum of a rows x R matrix: block partials + last block
__global__ void __launch_bounds__(WT) wide_colstat_kernel(const double *__restrict__ U, long long rows, int R, int first_iter,
                                                          double *colstat, double *partials, unsigned int *counter) {
    const int nblk = gridDim.x;
    const long long per = (rows + nblk - 1) / nblk;
    const long long r0 = blockIdx.x * per, r1 = min(rows, r0 + per);
    for (int c = threadIdx.x; c < R; c += WT) {
        double s = first_iter ? 0.0 : -DBL_MAX;
        for (long long i = r0; i < r1; ++i) {
            const double v = U[i * R + c];
            s = first_iter ? s + v * v : fmax(s, v);
        }
        partials[(size_t)blockIdx.x * R + c] = s;
    }
    if (last_block_ticket(counter, nblk)) {
        for (int c = threadIdx.x; c < R; c += WT) {
            double s = first_iter ? 0.0 : -DBL_MAX;
            for (int b = 0; b < nblk; ++b) {
                const double v = partials[(size_t)b * R + c];
                s = first_iter ? s + v : fmax(s, v);
            }
            colstat[c] = s;
        }
    }
}

// U <- U / lambda (lambda from the column statistic, cp.py:150-153) and ip = sum_r lambda_r sum_i U[i,r] Mip[i,r]
__global__ void __launch_bounds__(WT) wide_scale_kernel(double *U, const double *__restrict__ colstat, long long rows, int R,
                                                        int first_iter, double *lambda_out, const double *__restrict__ Mip,
                                                        double *ip_out, double *partials, unsigned int *counter) {
    __shared__ double red[WT / 32];
    const long long total = rows * R;
    double ip = 0.0;
    for (long long e = blockIdx.x * (long long)WT + threadIdx.x; e < total; e += (long long)gridDim.x * WT) {
        const int c = (int)(e % R);
        const double cs = colstat[c];
        const double l = first_iter ? sqrt(cs) : (cs < 1.0 ? 1.0 : cs);
        const double v = U[e] / l;
        U[e] = v;
        if (Mip) ip += l * v * Mip[e];
    }
    if (blockIdx.x == 0)
        for (int c = threadIdx.x; c < R; c += WT) {
            const double cs = colstat[c];
            lambda_out[c] = first_iter ? sqrt(cs) : (cs < 1.0 ? 1.0 : cs);
        }
    ip = warp_sum(ip);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ip;
    __syncthreads();
    if (threadIdx.x == 0) {
        double b = 0.0;
        for (int w = 0; w < WT / 32; ++w) b += red[w];
        partials[blockIdx.x] = b;
    }
    if (last_block_ticket(counter, gridDim.x)) {
        if (threadIdx.x == 0 && ip_out) {
            double s = 0.0;
            for (unsigned b = 0; b < gridDim.x; ++b) s += partials[b];
            ip_out[0] = s;
        }
    }
}

int wide_blocks(long long rows) { return (int)std::max<long long>(1, std::min<long long>((rows + 63) / 64, 2LL * sm_count())); }

}  // namespace

size_t wide_gram_ws(long long rows, int R) {
    const int64_t shp[2] = {(int64_t)std::max<long long>(rows, 1), R};
    return align_up(skt_unfold_gram_workspace(shp, 2, 1), 256);
}
size_t wide_pass_ws(long long rows, int R) { return 256 + align_up((size_t)wide_blocks(rows) * R * sizeof(double), 256); }

int wide_check_rank(int R) {
    SKT_REQUIRE(R >= 1 && R <= SKT_MAX_WIDE_RANK, SKT_EUNSUPPORTED, "rank %d outside [1, %d] supported by the solve kernels", R,
                SKT_MAX_WIDE_RANK);
    return SKT_OK;
}

int wide_gram(const double *U, long long rows, int R, double *G, void *ws, size_t ws_bytes, cudaStream_t st) {
    if (rows == 0) {
        SKT_CUDA(cudaMemsetAsync(G, 0, (size_t)R * R * sizeof(double), st));
        return SKT_OK;
    }
    const int64_t shp[2] = {(int64_t)rows, R};
    return skt_unfold_gram(U, shp, 2, 1, G, ws, ws_bytes, st);
}

int wide_hadamard_pinv(const double *G, int ndim, int R, int mode, double *P, int *info, cudaStream_t st) {
    const size_t eig_ws = skt_symeig_workspace(R, R);
    SKT_REQUIRE(eig_ws > 0, SKT_EUNSUPPORTED, "rank %d too large for the wide pseudo-inverse", R);
    const size_t mat = align_up((size_t)R * R * sizeof(double), 256);
    char *buf = nullptr;
    SKT_CUDA(cudaMallocAsync((void **)&buf, 2 * mat + align_up((size_t)R * sizeof(double), 256) + eig_ws, st));
    double *Y = (double *)buf, *V = (double *)(buf + mat), *W = (double *)(buf + 2 * mat);
    void *ews = buf + 2 * mat + align_up((size_t)R * sizeof(double), 256);
    SKT_CUDA(cudaMemsetAsync(info, 0, sizeof(int), st));
    wide_hadamard_kernel<<<(R * R + 255) / 256, 256, 0, st>>>(G, ndim, R, mode, Y, info);
    SKT_LAUNCH_CHECK();
    int rc = skt_symeig_top(Y, R, R, 0, W, V, nullptr, ews, eig_ws, st);
    if (rc == SKT_OK) {
        dim3 blk(16, 16), grd((R + 15) / 16, (R + 15) / 16);
        wide_pinv_from_eig_kernel<<<grd, blk, (size_t)R * sizeof(double), st>>>(V, W, R, P, info);
        ++g_launches;
        if (cudaGetLastError() != cudaSuccess) rc = SKT_ECUDA;
    }
    cudaFreeAsync(buf, st);
    return rc;
}

int wide_solve_stats(const double *M, const double *P, long long rows, int R, int first_iter, double *Unew, double *colstat,
                     void *ws, size_t ws_bytes, cudaStream_t st) {
    const size_t need = wide_pass_ws(rows, R);
    SKT_REQUIRE(ws && ws_bytes >= need, SKT_EWORKSPACE, "workspace too small: need %zu bytes, got %zu", need, ws_bytes);
    if (rows > 0) {
        const int64_t shp[2] = {(int64_t)rows, R};
        const int rc = skt_ttm_dense(M, shp, 2, P, R, 1, 0, Unew, st);      // Unew = M P^T = M P (P symmetric)
        if (rc) return rc;
    }
    unsigned int *counter = (unsigned int *)ws;
    double *partials = (double *)((char *)ws + 256);
    SKT_CUDA(cudaMemsetAsync(counter, 0, 256, st));
    wide_colstat_kernel<<<wide_blocks(rows), WT, 0, st>>>(Unew, rows, R, first_iter, colstat, partials, counter);
    SKT_LAUNCH_CHECK();
    return SKT_OK;
}

int wide_normalise_gram(double *U, const double *colstat, long long rows, int R, int first_iter, double *lambda, double *G,
                        const double *Mip, double *ip, void *ws, size_t ws_bytes, cudaStream_t st) {
    const size_t need = std::max(wide_pass_ws(rows, R), wide_gram_ws(rows, R));
    SKT_REQUIRE(ws && ws_bytes >= need, SKT_EWORKSPACE, "workspace too small: need %zu bytes, got %zu", need, ws_bytes);
    unsigned int *counter = (unsigned int *)ws;
    double *partials = (double *)((char *)ws + 256);
    SKT_CUDA(cudaMemsetAsync(counter, 0, 256, st));
    wide_scale_kernel<<<wide_blocks(rows), WT, 0, st>>>(U, colstat, rows, R, first_iter, lambda, Mip, ip, partials, counter);
    SKT_LAUNCH_CHECK();
    return wide_gram(U, rows, R, G, ws, ws_bytes, st);
}

}  // namespace skt

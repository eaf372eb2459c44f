// CP-ALS normal equations for sm_100a: tall-skinny Grams, Hadamard-of-Grams pseudo-inverse,
// factor update + normalisation, Kruskal fit.  Reference: sktensor/cp.py:144-159 and
// sktensor/ktensor.py:113-150.  Every reduction is ordered (partials + "last block reduces"),
// so results are bitwise reproducible run to run.
#include <algorithm>
#include <cfloat>

#include "common.cuh"
#include "pinv.cuh"

namespace skt {
// rank > SKT_MAX_RANK: the generic path of cpals_wide.cu behind the same entry points
size_t wide_gram_ws(long long rows, int R);
size_t wide_pass_ws(long long rows, int R);
int wide_check_rank(int R);
int wide_gram(const double *U, long long rows, int R, double *G, void *ws, size_t ws_bytes, cudaStream_t st);
int wide_hadamard_pinv(const double *G, int ndim, int R, int mode, double *P, int *info, cudaStream_t st);
int wide_solve_stats(const double *M, const double *P, long long rows, int R, int first_iter, double *Unew, double *colstat,
                     void *ws, size_t ws_bytes, cudaStream_t st);
int wide_normalise_gram(double *U, const double *colstat, long long rows, int R, int first_iter, double *lambda, double *G,
                        const double *Mip, double *ip, void *ws, size_t ws_bytes, cudaStream_t st);

namespace {

constexpr int ROWS = 64;        // factor rows staged per tile
constexpr int NT = 256;         // 16 x 16 threads
constexpr int MAXR = SKT_MAX_RANK;

// Deterministic block-wide sum / max of one double per thread (NT threads). Result valid in thread 0.
__device__ __forceinline__ double block_sum(double v, double *red) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    __syncthreads();
    return s;
}

// ---------------------------------------------------------------------------------------------
// Factor pass:  (optionally) U <- U / lambda,  G = U^T U,  ip = sum_r lambda_r sum_i U[i,r] Mip[i,r]
// 16x16 threads; thread (ty,tx) owns G[ty+16a][tx+16b], a,b < TT = ceil(R/16).
// ---------------------------------------------------------------------------------------------
template <int TT>
__global__ void __launch_bounds__(NT) factor_pass_kernel(double *U, const double *__restrict__ colstat, long long rows,
                                                         int R, int normalise, int first_iter, double *lambda_out,
                                                         double *G_out, const double *__restrict__ Mip, double *ip_out,
                                                         double *partials, unsigned int *counter) {
    __shared__ double Us[ROWS][16 * TT + 1];
    __shared__ double lam[16 * TT];
    __shared__ double red[NT / 32];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int nblk = gridDim.x;

    if (tid < 16 * TT) {
        double l = 1.0;
        if (normalise && tid < R) {
            const double c = colstat[tid];
            l = first_iter ? sqrt(c) : (c < 1.0 ? 1.0 : c);  // cp.py:150 / cp.py:152-153
        }
        lam[tid] = l;
        if (normalise && lambda_out && blockIdx.x == 0 && tid < R) lambda_out[tid] = l;
    }
    for (int e = tid; e < ROWS * (16 * TT + 1); e += NT) (&Us[0][0])[e] = 0.0;
    __syncthreads();

    double g[TT][TT];
#pragma unroll
    for (int a = 0; a < TT; ++a)
#pragma unroll
        for (int b = 0; b < TT; ++b) g[a][b] = 0.0;
    double ip = 0.0;

    const long long ntile = (rows + ROWS - 1) / ROWS;
    for (long long tile = blockIdx.x; tile < ntile; tile += nblk) {
        const long long r0 = tile * ROWS;
        const int nr = (int)min((long long)ROWS, rows - r0);
        // coalesced load of nr x R, divide, write back
        for (int e = tid; e < ROWS * R; e += NT) {
            const int i = e / R, c = e % R;
            double v = 0.0;
            if (i < nr) {
                const long long gi = (r0 + i) * R + c;
                v = U[gi];
                if (normalise) {
                    v = v / lam[c];
                    U[gi] = v;
                }
                if (Mip) ip += lam[c] * v * Mip[gi];
            }
            Us[i][c] = v;
        }
        __syncthreads();
        for (int i = 0; i < nr; ++i) {
            double ua[TT], ub[TT];
#pragma unroll
            for (int a = 0; a < TT; ++a) ua[a] = Us[i][ty + 16 * a];
#pragma unroll
            for (int b = 0; b < TT; ++b) ub[b] = Us[i][tx + 16 * b];
#pragma unroll
            for (int a = 0; a < TT; ++a)
#pragma unroll
                for (int b = 0; b < TT; ++b) g[a][b] += ua[a] * ub[b];
        }
        __syncthreads();
    }
    double *mine = partials + (size_t)blockIdx.x * (MAXR * MAXR + 1);
#pragma unroll
    for (int a = 0; a < TT; ++a)
#pragma unroll
        for (int b = 0; b < TT; ++b) {
            const int ra = ty + 16 * a, cb = tx + 16 * b;
            if (ra < R && cb < R) mine[ra * R + cb] = g[a][b];
        }
    const double ipb = block_sum(ip, red);
    if (tid == 0) mine[MAXR * MAXR] = ipb;

    if (last_block_ticket(counter, nblk)) {
        for (int e = tid; e < R * R; e += NT) {
            double s = 0.0;
            for (int b = 0; b < nblk; ++b) s += partials[(size_t)b * (MAXR * MAXR + 1) + e];
            G_out[e] = s;
        }
        if (tid == 0 && ip_out) {
            double s = 0.0;
            for (int b = 0; b < nblk; ++b) s += partials[(size_t)b * (MAXR * MAXR + 1) + MAXR * MAXR];
            ip_out[0] = s;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Unew = M P  and per-column statistic (sum of squares or signed max).  cp.py:147-152
// ---------------------------------------------------------------------------------------------
template <int TT>
__global__ void __launch_bounds__(NT) solve_stats_kernel(const double *__restrict__ M, const double *__restrict__ P,
                                                         long long rows, int R, int first_iter, double *Unew,
                                                         double *colstat, double *partials, unsigned int *counter) {
    extern __shared__ double sm[];
    double *Ps = sm;                         // [R][16*TT+1]
    double *Ms = Ps + MAXR * (16 * TT + 1);  // [ROWS][R+1]
    double *cs = Ms + ROWS * (MAXR + 1);     // [16][16*TT]
    const int PL = 16 * TT + 1, ML = R + 1;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int nblk = gridDim.x;

    for (int e = tid; e < R * 16 * TT; e += NT) {
        const int k = e / (16 * TT), c = e % (16 * TT);
        Ps[k * PL + c] = (c < R) ? P[k * R + c] : 0.0;
    }
    double stat[TT];
#pragma unroll
    for (int b = 0; b < TT; ++b) stat[b] = first_iter ? 0.0 : -DBL_MAX;
    __syncthreads();

    const long long ntile = (rows + ROWS - 1) / ROWS;
    for (long long tile = blockIdx.x; tile < ntile; tile += nblk) {
        const long long r0 = tile * ROWS;
        const int nr = (int)min((long long)ROWS, rows - r0);
        for (int e = tid; e < ROWS * R; e += NT) {
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
            if (i < nr) {
#pragma unroll
                for (int b = 0; b < TT; ++b) {
                    const int c = tx + 16 * b;
                    if (c < R) {
                        Unew[(r0 + i) * R + c] = acc[a][b];
                        if (first_iter) stat[b] += acc[a][b] * acc[a][b];
                        else stat[b] = nanmax(stat[b], acc[a][b]);
                    }
                }
            }
        }
        __syncthreads();
    }
    // combine the 16 row-groups (ty) per column in a fixed order
#pragma unroll
    for (int b = 0; b < TT; ++b) cs[ty * 16 * TT + tx + 16 * b] = stat[b];
    __syncthreads();
    double *mine = partials + (size_t)blockIdx.x * MAXR;
    if (tid < R) {
        double s = first_iter ? 0.0 : -DBL_MAX;
        for (int y = 0; y < 16; ++y) {
            const double v = cs[y * 16 * TT + tid];
            s = first_iter ? s + v : nanmax(s, v);
        }
        mine[tid] = s;
    }
    if (last_block_ticket(counter, nblk)) {
        if (tid < R) {
            double s = first_iter ? 0.0 : -DBL_MAX;
            for (int b = 0; b < nblk; ++b) {
                const double v = partials[(size_t)b * MAXR + tid];
                s = first_iter ? s + v : nanmax(s, v);
            }
            colstat[tid] = s;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// P = pinv( hadamard_{m != mode} G_m ),  scipy.linalg.pinv semantics (cp.py:144-147); one CTA.
// The algorithm lives in pinv.cuh (shared with the fused small-factor update).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) hadamard_pinv_kernel(const double *__restrict__ G, int ndim, int R, int mode,
                                                             double *P, int *info) {
    extern __shared__ double jsm[];
    const PinvShared s = pinv_carve(jsm, R);
    if (pinv_load_y(s, G, ndim, R, mode)) {
        for (int e = threadIdx.x; e < R * R; e += blockDim.x) P[e] = nan("");
        if (threadIdx.x == 0) info[0] = -1;
        return;
    }
    const int kept = block_pinv(s, G, ndim, R, mode, [&](int a, int b, double v) { P[a * R + b] = v; });
    if (threadIdx.x == 0) info[0] = kept;
}

// fit (cp.py:157-159) with ||P||^2 from the Grams (ktensor.py:113-126)
__global__ void cp_fit_kernel(const double *__restrict__ G, int ndim, int R, const double *__restrict__ lambda,
                              const double *__restrict__ ip, const double *__restrict__ normX2, double *out) {
    __shared__ double red[NT / 32];
    double s = 0.0;
    for (int e = threadIdx.x; e < R * R; e += NT) {
        const int a = e / R, b = e % R;
        double y = lambda[a] * lambda[b];
        for (int k = 0; k < ndim; ++k) y *= G[(size_t)k * R * R + e];
        s += y;
    }
    const double normP2 = block_sum(s, red);
    if (threadIdx.x == 0) {
        const double nx2 = normX2[0];
        out[0] = 1.0 - (nx2 + normP2 - 2.0 * ip[0]) / nx2;
        out[1] = normP2;
    }
}

__global__ void __launch_bounds__(NT) sumsq_kernel(const double *__restrict__ x, long long n, double *out, double *partials,
                                                   unsigned int *counter) {
    __shared__ double red[NT / 32];
    double s = 0.0;
    for (long long i = blockIdx.x * (long long)NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
        const double v = x[i];
        s += v * v;
    }
    const double b = block_sum(s, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = b;
    if (last_block_ticket(counter, gridDim.x)) {
        double t = 0.0;
        for (int k = threadIdx.x; k < (int)gridDim.x; k += NT) t += partials[k];
        const double tot = block_sum(t, red);
        if (threadIdx.x == 0) out[0] = tot;
    }
}

// out = sum_i sum_r lambda_r U[i,r] M[i,r]   (<P,X> from the last-mode MTTKRP; ktensor.py:128-150)
__global__ void __launch_bounds__(NT) weighted_dot_kernel(const double *__restrict__ U, const double *__restrict__ M,
                                                          const double *__restrict__ lambda, long long n, int R,
                                                          double *out, double *partials, unsigned int *counter) {
    __shared__ double red[NT / 32];
    double s = 0.0;
    for (long long i = blockIdx.x * (long long)NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
        const double l = lambda ? lambda[i % R] : 1.0;
        s += l * U[i] * M[i];
    }
    const double b = block_sum(s, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = b;
    if (last_block_ticket(counter, gridDim.x)) {
        double t = 0.0;
        for (int k = threadIdx.x; k < (int)gridDim.x; k += NT) t += partials[k];
        const double tot = block_sum(t, red);
        if (threadIdx.x == 0) out[0] = tot;
    }
}

int pass_blocks(long long rows) {
    const long long ntile = (rows + ROWS - 1) / ROWS;
    return (int)std::max<long long>(1, std::min<long long>(ntile, 2LL * sm_count()));
}

// workspace layout shared by the factor passes: [counter (256 B)] [partials]
size_t pass_ws_bytes(long long rows, size_t per_block_doubles) {
    return 256 + align_up((size_t)pass_blocks(rows) * per_block_doubles * sizeof(double), 256);
}

int check_rank(int rank) {
    SKT_REQUIRE(rank >= 1 && rank <= MAXR, SKT_EUNSUPPORTED, "rank %d outside [1, %d] supported by the solve kernels", rank, MAXR);
    return SKT_OK;
}

int launch_factor_pass(double *U, const double *colstat, long long rows, int R, int normalise, int first_iter,
                       double *lambda, double *G, const double *Mip, double *ip, void *ws, size_t ws_bytes,
                       cudaStream_t st) {
    const size_t need = pass_ws_bytes(rows, MAXR * MAXR + 1);
    SKT_REQUIRE(ws && ws_bytes >= need, SKT_EWORKSPACE, "workspace too small: need %zu bytes, got %zu", need, ws_bytes);
    unsigned int *counter = (unsigned int *)ws;
    double *partials = (double *)((char *)ws + 256);
    SKT_CUDA(cudaMemsetAsync(counter, 0, 256, st));
    const int nb = pass_blocks(rows);
    const int tt = (R + 15) / 16;
#define FP(TT_) factor_pass_kernel<TT_><<<nb, NT, 0, st>>>(U, colstat, rows, R, normalise, first_iter, lambda, G, Mip, ip, partials, counter)
    switch (tt) {
        case 1: FP(1); break;
        case 2: FP(2); break;
        case 3: FP(3); break;
        default: FP(4); break;
    }
#undef FP
    SKT_LAUNCH_CHECK();
    return SKT_OK;
}

}  // namespace
}  // namespace skt

using namespace skt;

extern "C" size_t skt_gram_workspace(int64_t rows, int rank) {
    return rank > MAXR ? wide_gram_ws(rows, rank) : pass_ws_bytes(rows, MAXR * MAXR + 1);
}
extern "C" size_t skt_normalise_gram_workspace(int64_t rows, int rank) {
    return rank > MAXR ? std::max(wide_gram_ws(rows, rank), wide_pass_ws(rows, rank)) : pass_ws_bytes(rows, MAXR * MAXR + 1);
}
extern "C" size_t skt_solve_stats_workspace(int64_t rows, int rank) {
    return rank > MAXR ? wide_pass_ws(rows, rank) : pass_ws_bytes(rows, MAXR);
}
extern "C" size_t skt_sumsq_workspace(int64_t n) {
    (void)n;
    return 256 + align_up((size_t)4 * sm_count() * sizeof(double), 256);
}

extern "C" int skt_gram(const double *U_d, int64_t rows, int rank, double *G_d, void *ws_d, size_t ws_bytes,
                        skt_stream stream) {
    SKT_REQUIRE((U_d || rows == 0) && G_d && rows >= 0, SKT_EINVAL, "NULL argument or negative rows");
    if (rank > MAXR) {
        const int rcw = wide_check_rank(rank);
        return rcw ? rcw : wide_gram(U_d, rows, rank, G_d, ws_d, ws_bytes, (cudaStream_t)stream);
    }
    int rc = check_rank(rank);
    if (rc) return rc;
    return launch_factor_pass(const_cast<double *>(U_d), nullptr, rows, rank, 0, 0, nullptr, G_d, nullptr, nullptr, ws_d,
                              ws_bytes, (cudaStream_t)stream);
}

extern "C" int skt_normalise_gram(double *U_d, const double *colstat_d, int64_t rows, int rank, int first_iter,
                                  double *lambda_d, double *G_d, const double *Mip_d, double *ip_d, void *ws_d,
                                  size_t ws_bytes, skt_stream stream) {
    SKT_REQUIRE((U_d || rows == 0) && colstat_d && lambda_d && G_d && rows >= 0, SKT_EINVAL, "NULL argument or negative rows");
    SKT_REQUIRE((Mip_d == nullptr) == (ip_d == nullptr) || rows == 0, SKT_EINVAL, "Mip_d and ip_d must be given together");
    if (rank > MAXR) {
        const int rcw = wide_check_rank(rank);
        return rcw ? rcw : wide_normalise_gram(U_d, colstat_d, rows, rank, first_iter, lambda_d, G_d, Mip_d, ip_d, ws_d, ws_bytes,
                                               (cudaStream_t)stream);
    }
    int rc = check_rank(rank);
    if (rc) return rc;
    return launch_factor_pass(U_d, colstat_d, rows, rank, 1, first_iter, lambda_d, G_d, Mip_d, ip_d, ws_d, ws_bytes,
                              (cudaStream_t)stream);
}

extern "C" int skt_solve_stats(const double *M_d, const double *P_d, int64_t rows, int rank, int first_iter,
                               double *Unew_d, double *colstat_d, void *ws_d, size_t ws_bytes, skt_stream stream) {
    SKT_REQUIRE((M_d || rows == 0) && P_d && (Unew_d || rows == 0) && colstat_d && rows >= 0, SKT_EINVAL, "NULL argument or negative rows");
    if (rank > MAXR) {
        const int rcw = wide_check_rank(rank);
        return rcw ? rcw : wide_solve_stats(M_d, P_d, rows, rank, first_iter, Unew_d, colstat_d, ws_d, ws_bytes, (cudaStream_t)stream);
    }
    int rc = check_rank(rank);
    if (rc) return rc;
    const size_t need = pass_ws_bytes(rows, MAXR);
    SKT_REQUIRE(ws_d && ws_bytes >= need, SKT_EWORKSPACE, "workspace too small: need %zu bytes, got %zu", need, ws_bytes);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned int *counter = (unsigned int *)ws_d;
    double *partials = (double *)((char *)ws_d + 256);
    SKT_CUDA(cudaMemsetAsync(counter, 0, 256, st));
    const int nb = pass_blocks(rows);
    const int tt = (rank + 15) / 16;
    const size_t smem = (size_t)(MAXR * (16 * tt + 1) + ROWS * (MAXR + 1) + 16 * 16 * tt) * sizeof(double);
#define SS(TT_)                                                                                                   \
    do {                                                                                                          \
        static unsigned long long cfg = 0;   /* bit d: configured on device d */                                                                                  \
        if (skt::first_use_on_device(cfg)) {                                                                                               \
            SKT_CUDA(cudaFuncSetAttribute(solve_stats_kernel<TT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        }                                                                                                         \
        solve_stats_kernel<TT_><<<nb, NT, smem, st>>>(M_d, P_d, rows, rank, first_iter, Unew_d, colstat_d, partials, counter); \
    } while (0)
    switch (tt) {
        case 1: SS(1); break;
        case 2: SS(2); break;
        case 3: SS(3); break;
        default: SS(4); break;
    }
#undef SS
    SKT_LAUNCH_CHECK();
    return SKT_OK;
}

extern "C" int skt_hadamard_pinv(const double *G_d, int ndim, int rank, int mode, double *P_d, int *info_d,
                                 skt_stream stream) {
    SKT_REQUIRE(G_d && P_d && info_d, SKT_EINVAL, "NULL argument");
    SKT_REQUIRE(ndim >= 1 && ndim <= SKT_MAX_NDIM && mode >= -1 && mode < ndim, SKT_EINVAL, "bad ndim/mode");
    if (rank > MAXR) {
        const int rcw = wide_check_rank(rank);
        return rcw ? rcw : wide_hadamard_pinv(G_d, ndim, rank, mode, P_d, info_d, (cudaStream_t)stream);
    }
    int rc = check_rank(rank);
    if (rc) return rc;
    const size_t smem = pinv_smem_doubles(rank) * sizeof(double);
    static unsigned long long cfg = 0;   /* bit d: configured on device d */
    if (skt::first_use_on_device(cfg)) {
        SKT_CUDA(cudaFuncSetAttribute(hadamard_pinv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(pinv_smem_doubles(MAXR) * sizeof(double))));
    }
    hadamard_pinv_kernel<<<1, 1024, smem, (cudaStream_t)stream>>>(G_d, ndim, rank, mode, P_d, info_d);
    SKT_LAUNCH_CHECK();
    return SKT_OK;
}

extern "C" int skt_cp_fit(const double *G_d, int ndim, int rank, const double *lambda_d, const double *ip_d,
                          const double *normX2_d, double *out_d, skt_stream stream) {
    SKT_REQUIRE(G_d && lambda_d && ip_d && normX2_d && out_d, SKT_EINVAL, "NULL argument");
    SKT_REQUIRE(ndim >= 1 && ndim <= SKT_MAX_NDIM, SKT_EINVAL, "bad ndim");
    int rc = wide_check_rank(rank);          // the fit kernel loops over R x R: any rank the solve path accepts
    if (rc) return rc;
    cp_fit_kernel<<<1, NT, 0, (cudaStream_t)stream>>>(G_d, ndim, rank, lambda_d, ip_d, normX2_d, out_d);
    SKT_LAUNCH_CHECK();
    return SKT_OK;
}

extern "C" int skt_sumsq(const double *x_d, int64_t n, double *out_d, void *ws_d, size_t ws_bytes, skt_stream stream) {
    SKT_REQUIRE((x_d || n == 0) && out_d && n >= 0, SKT_EINVAL, "NULL argument or negative length");
    const size_t need = skt_sumsq_workspace(n);
    SKT_REQUIRE(ws_d && ws_bytes >= need, SKT_EWORKSPACE, "workspace too small: need %zu bytes, got %zu", need, ws_bytes);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned int *counter = (unsigned int *)ws_d;
    double *partials = (double *)((char *)ws_d + 256);
    SKT_CUDA(cudaMemsetAsync(counter, 0, 256, st));
    const int nb = (int)std::max<long long>(1, std::min<long long>((n + NT - 1) / NT, 4LL * sm_count()));
    sumsq_kernel<<<nb, NT, 0, st>>>(x_d, n, out_d, partials, counter);
    SKT_LAUNCH_CHECK();
    return SKT_OK;
}

extern "C" int skt_weighted_dot(const double *U_d, const double *M_d, const double *lambda_d, int64_t rows, int rank,
                                double *out_d, void *ws_d, size_t ws_bytes, skt_stream stream) {
    SKT_REQUIRE(((U_d && M_d) || rows == 0) && out_d && rows >= 0 && rank >= 1, SKT_EINVAL, "NULL argument or bad size");
    const long long n = (long long)rows * rank;
    const size_t need = skt_sumsq_workspace(n);
    SKT_REQUIRE(ws_d && ws_bytes >= need, SKT_EWORKSPACE, "workspace too small: need %zu bytes, got %zu", need, ws_bytes);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned int *counter = (unsigned int *)ws_d;
    double *partials = (double *)((char *)ws_d + 256);
    SKT_CUDA(cudaMemsetAsync(counter, 0, 256, st));
    const int nb = (int)std::max<long long>(1, std::min<long long>((n + NT - 1) / NT, 4LL * sm_count()));
    weighted_dot_kernel<<<nb, NT, 0, st>>>(U_d, M_d, lambda_d, n, rank, out_d, partials, counter);
    SKT_LAUNCH_CHECK();
    return SKT_OK;
}

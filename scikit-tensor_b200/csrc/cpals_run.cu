// The whole CP-ALS loop of cp.als (reference: sktensor/cp.py:132-171) behind ONE C call, for problems whose factors
// fit the single-launch update kernel: per mode one MTTKRP launch and one fused update launch (pinv, solve, normalise,
// Gram; the last mode also <P,X> and the fit), per iteration one 24-byte read-back for the convergence test.
// Small tensors are pure launch latency (BASELINE config 1: 880 elements, 22 iterations); driving ~6 launches per
// iteration from C instead of ~14 through the Python engine is what this entry point is for.
#include <algorithm>
#include <chrono>
#include <cmath>

#include "common.cuh"

using namespace skt;

namespace {
struct RunLayout {
    size_t off_scal, off_info, off_G, off_M, off_mt, total;
    size_t mt_bytes;
};
int run_layout(const int64_t *shape, int ndim, int rank, RunLayout &L) {
    long long maxrows = 0;
    size_t mt = skt_sumsq_workspace(1);
    for (int m = 0; m < ndim; ++m) {
        maxrows = std::max<long long>(maxrows, shape[m]);
        const size_t w = skt_mttkrp_dense_workspace(shape, ndim, rank, m);
        SKT_REQUIRE(w > 0, SKT_EUNSUPPORTED, "no MTTKRP plan for mode %d", m);
        mt = std::max(mt, w);
        mt = std::max(mt, skt_gram_workspace(shape[m], rank));
    }
    size_t o = 0;
    L.off_scal = o; o += 256;                                         // normX2, ip, fit[2]
    L.off_info = o; o += align_up((size_t)ndim * sizeof(int), 256);
    L.off_G = o; o += align_up((size_t)ndim * rank * rank * sizeof(double), 256);
    L.off_M = o; o += align_up((size_t)maxrows * rank * sizeof(double), 256);
    L.off_mt = o; L.mt_bytes = align_up(mt, 256); o += L.mt_bytes;
    L.total = o;
    return SKT_OK;
}
}  // namespace

extern "C" int skt_cp_als_run_fits(const int64_t *shape, int ndim, int rank) {
    if (!shape || ndim < 2 || ndim > SKT_MAX_NDIM || rank < 1 || rank > SKT_MAX_RANK) return 0;
    for (int m = 0; m < ndim; ++m)
        if (shape[m] < 1 || !skt_cp_update_fused_fits(shape[m], rank)) return 0;
    return 1;
}

extern "C" size_t skt_cp_als_run_workspace(const int64_t *shape, int ndim, int rank) {
    RunLayout L;
    if (!skt_cp_als_run_fits(shape, ndim, rank) || run_layout(shape, ndim, rank, L) != SKT_OK) return 0;
    return L.total;
}

extern "C" int skt_cp_als_run(const double *X_d, const int64_t *shape, int ndim, int rank, double *const *U_d,
                              unsigned init_mask, int max_iter, double conv, int fit_full, double *lambda_d, void *ws_d,
                              size_t ws_bytes, double *fit_h, int *itr_h, double *exectimes_h, skt_stream stream) {
    SKT_REQUIRE(X_d && shape && U_d && lambda_d && fit_h && itr_h, SKT_EINVAL, "NULL argument");
    SKT_REQUIRE(skt_cp_als_run_fits(shape, ndim, rank), SKT_EUNSUPPORTED, "the factors do not fit the single-launch update kernel");
    RunLayout L;
    int rc = run_layout(shape, ndim, rank, L);
    if (rc) return rc;
    SKT_REQUIRE(ws_d && ws_bytes >= L.total, SKT_EWORKSPACE, "workspace too small: need %zu bytes, got %zu", L.total, ws_bytes);
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = (char *)ws_d;
    double *normX2 = (double *)(ws + L.off_scal), *ip = normX2 + 1, *fitd = normX2 + 2;
    int *info = (int *)(ws + L.off_info);
    double *G = (double *)(ws + L.off_G), *M = (double *)(ws + L.off_M);
    void *mt = ws + L.off_mt;
    long long total = 1;
    for (int m = 0; m < ndim; ++m) {
        SKT_REQUIRE(U_d[m] != nullptr, SKT_EINVAL, "U[%d] is NULL (pass storage for every factor)", m);
        total *= shape[m];
    }
    SKT_CUDA(cudaMemsetAsync(ws + L.off_scal, 0, 256, st));
    SKT_CUDA(cudaMemsetAsync(info, 0, (size_t)ndim * sizeof(int), st));
    SKT_CUDA(cudaMemsetAsync(G, 0, (size_t)ndim * rank * rank * sizeof(double), st));
    rc = skt_sumsq(X_d, total, normX2, mt, L.mt_bytes, st);
    if (rc) return rc;
    for (int m = 0; m < ndim; ++m)
        if (init_mask & (1u << m)) {
            rc = skt_gram(U_d[m], shape[m], rank, G + (size_t)m * rank * rank, mt, L.mt_bytes, st);
            if (rc) return rc;
        }
    const double *Uc[SKT_MAX_NDIM];
    for (int m = 0; m < ndim; ++m) Uc[m] = U_d[m];
    double fit = 0.0;
    int itr = -1;
    struct { double fit[2]; } hfit;
    int hinfo[SKT_MAX_NDIM];
    for (itr = 0; itr < max_iter; ++itr) {
        const auto tic = std::chrono::steady_clock::now();
        const double fitold = fit;
        for (int n = 0; n < ndim; ++n) {
            rc = skt_mttkrp_dense(X_d, shape, ndim, Uc, rank, n, M, mt, L.mt_bytes, st);     // U[n] is ignored (cp.py:143)
            if (rc) return rc;
            const bool last = fit_full && n == ndim - 1;
            rc = skt_cp_update_fused(M, shape[n], rank, ndim, n, itr == 0 ? 1 : 0, G, U_d[n], lambda_d, info + n,
                                     last ? ip : nullptr, last ? normX2 : nullptr, last ? fitd : nullptr, st);
            if (rc) return rc;
        }
        SKT_CUDA(cudaMemcpyAsync(&hfit, fitd, sizeof(hfit), cudaMemcpyDeviceToHost, st));
        SKT_CUDA(cudaMemcpyAsync(hinfo, info, (size_t)ndim * sizeof(int), cudaMemcpyDeviceToHost, st));
        SKT_CUDA(cudaStreamSynchronize(st));
        for (int n = 0; n < ndim; ++n)
            if (hinfo[n] < 0) {                      // scipy.linalg.pinv raises on a non-finite Gram product (cp.py:147)
                set_error("array must not contain infs or NaNs");
                *itr_h = itr;
                return SKT_ENONFINITE;
            }
        fit = fit_full ? hfit.fit[0] : (double)itr;  // cp.py:157-161
        if (exectimes_h)
            exectimes_h[itr] = std::chrono::duration<double>(std::chrono::steady_clock::now() - tic).count();
        if (itr > 0 && std::fabs(fitold - fit) < conv) { ++itr; break; }
    }
    *fit_h = fit;
    *itr_h = itr - 1;                                // index of the last iteration, like the reference's loop variable
    return SKT_OK;
}

// Tensor-times-matrix and unfolding Gram for the HOOI path (reference: dtensor._ttm_compute,
// sktensor/dtensor.py:58-76, and the Gram inside core.nvecs, sktensor/core.py:279,285).
//
// Both are expressed on the free (L, I_n, T) view of a C-order tensor -- no transpose/reshape
// copies as in the reference -- and run through one strided, batched, shared-memory tiled FP64
// GEMM:   C[b](m,n) (+)= sum_k A[b](m,k) * B[b](k,n).
#include <algorithm>

#include "common.cuh"

namespace skt {
namespace {

constexpr int TM = 64, TN = 64, TK = 16, GT = 256;

struct GemmParams {
    const double *A, *B;
    double *C;
    long long am, ak, ab;  // element strides of A: row, k, batch
    long long bk, bn, bb;
    long long cm, cn, cb;
    int M, N, K;
    int nbatch;
    int sum_batches;  // 1: C = sum_b A[b] B[b]  (single output), 0: one output per batch
};

__global__ void __launch_bounds__(GT) gemm_strided_kernel(const __grid_constant__ GemmParams p) {
    __shared__ double As[TK][TM + 1];
    __shared__ double Bs[TK][TN + 1];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    const int b_lo = p.sum_batches ? 0 : blockIdx.z;
    const int b_hi = p.sum_batches ? p.nbatch : blockIdx.z + 1;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    // pick the thread->element mapping of the tile loads so that the unit-stride axis is the fast one
    const bool a_kfast = (p.ak == 1), b_kfast = (p.bk == 1);

    for (int b = b_lo; b < b_hi; ++b) {
        const double *A = p.A + (long long)b * p.ab;
        const double *B = p.B + (long long)b * p.bb;
        for (int k0 = 0; k0 < p.K; k0 += TK) {
            for (int e = tid; e < TM * TK; e += GT) {
                const int kk = a_kfast ? (e % TK) : (e / TM);
                const int mm = a_kfast ? (e / TK) : (e % TM);
                const int gm = m0 + mm, gk = k0 + kk;
                As[kk][mm] = (gm < p.M && gk < p.K) ? A[gm * p.am + gk * p.ak] : 0.0;
            }
            for (int e = tid; e < TN * TK; e += GT) {
                const int kk = b_kfast ? (e % TK) : (e / TN);
                const int nn = b_kfast ? (e / TK) : (e % TN);
                const int gn = n0 + nn, gk = k0 + kk;
                Bs[kk][nn] = (gn < p.N && gk < p.K) ? B[gk * p.bk + gn * p.bn] : 0.0;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < TK; ++kk) {
                double a[4], bv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = As[kk][ty + 16 * i];
#pragma unroll
                for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx + 16 * j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * bv[j];
            }
            __syncthreads();
        }
    }
    double *C = p.C + (p.sum_batches ? 0 : (long long)blockIdx.z * p.cb);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gm = m0 + ty + 16 * i, gn = n0 + tx + 16 * j;
            if (gm < p.M && gn < p.N) C[gm * p.cm + gn * p.cn] = acc[i][j];
        }
}

int launch_gemm(const GemmParams &p, cudaStream_t st) {
    const int gz = p.sum_batches ? 1 : p.nbatch;
    SKT_REQUIRE(gz <= 65535, SKT_ESHAPE, "too many batches (%d) for the TTM kernel", gz);
    dim3 grid((p.N + TN - 1) / TN, (p.M + TM - 1) / TM, gz);
    SKT_REQUIRE(grid.y <= 65535, SKT_ESHAPE, "TTM output too tall");
    gemm_strided_kernel<<<grid, GT, 0, st>>>(p);
    SKT_LAUNCH_CHECK();
    return SKT_OK;
}

int view3(const int64_t *shape, int ndim, int mode, long long &L, long long &In, long long &T) {
    SKT_REQUIRE(shape && ndim >= 1 && ndim <= SKT_MAX_NDIM && mode >= 0 && mode < ndim, SKT_EINVAL, "bad shape/ndim/mode");
    L = 1; T = 1;
    for (int m = 0; m < mode; ++m) L *= shape[m];
    for (int m = mode + 1; m < ndim; ++m) T *= shape[m];
    In = shape[mode];
    SKT_REQUIRE(L < (1LL << 31) && T < (1LL << 31) && In < (1LL << 31), SKT_ESHAPE, "tensor view too large");
    return SKT_OK;
}

}  // namespace
}  // namespace skt

using namespace skt;

extern "C" int skt_ttm_dense(const double *X_d, const int64_t *shape, int ndim, const double *V_d, int64_t pdim,
                             int mode, int transp, double *Y_d, skt_stream stream) {
    SKT_REQUIRE(X_d && V_d && Y_d && pdim >= 1, SKT_EINVAL, "NULL argument or bad output size");
    long long L, In, T;
    int rc = view3(shape, ndim, mode, L, In, T);
    if (rc) return rc;
    GemmParams p;
    p.K = (int)In;
    p.sum_batches = 0;
    // A(j, i) = V^T or V
    const long long a_row = transp ? 1 : In, a_k = transp ? pdim : 1;
    if (T > 1) {
        // per l:  Y_l (p x T) = A (p x I_n) . X_l (I_n x T)
        p.A = V_d; p.am = a_row; p.ak = a_k; p.ab = 0;
        p.B = X_d; p.bk = T; p.bn = 1; p.bb = In * T;
        p.C = Y_d; p.cm = T; p.cn = 1; p.cb = pdim * T;
        p.M = (int)pdim; p.N = (int)T; p.nbatch = (int)L;
        if (L > 65535) {
            // fold batches into rows is impossible here (strides differ); run in slabs
            for (long long l0 = 0; l0 < L; l0 += 65535) {
                GemmParams q = p;
                q.nbatch = (int)std::min<long long>(65535, L - l0);
                q.B = X_d + l0 * In * T;
                q.C = Y_d + l0 * pdim * T;
                rc = launch_gemm(q, (cudaStream_t)stream);
                if (rc) return rc;
            }
            return SKT_OK;
        }
    } else {
        // last mode:  Y (L x p) = X (L x I_n) . A^T (I_n x p)
        p.A = X_d; p.am = In; p.ak = 1; p.ab = 0;
        p.B = V_d; p.bk = a_k; p.bn = a_row; p.bb = 0;
        p.C = Y_d; p.cm = pdim; p.cn = 1; p.cb = 0;
        p.M = (int)L; p.N = (int)pdim; p.nbatch = 1;
    }
    return launch_gemm(p, (cudaStream_t)stream);
}

extern "C" size_t skt_unfold_gram_workspace(const int64_t *shape, int ndim, int mode) {
    (void)shape; (void)ndim; (void)mode;
    return 256;
}

extern "C" int skt_unfold_gram(const double *X_d, const int64_t *shape, int ndim, int mode, double *G_d, void *ws_d,
                               size_t ws_bytes, skt_stream stream) {
    (void)ws_d; (void)ws_bytes;
    SKT_REQUIRE(X_d && G_d, SKT_EINVAL, "NULL argument");
    long long L, In, T;
    int rc = view3(shape, ndim, mode, L, In, T);
    if (rc) return rc;
    // G = sum_l X_l X_l^T  with X_l (I_n x T):  A(i,t) = X_l[i,t], B(t,i') = X_l[i',t]
    GemmParams p;
    p.A = X_d; p.am = T; p.ak = 1; p.ab = In * T;
    p.B = X_d; p.bk = 1; p.bn = T; p.bb = In * T;
    p.C = G_d; p.cm = In; p.cn = 1; p.cb = 0;
    p.M = (int)In; p.N = (int)In; p.K = (int)T; p.nbatch = (int)L; p.sum_batches = 1;
    return launch_gemm(p, (cudaStream_t)stream);
}

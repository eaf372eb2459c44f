// Tensor-times-matrix and unfolding Gram for the HOOI path (reference: dtensor._ttm_compute,
// sktensor/dtensor.py:58-76, and the Gram inside core.nvecs, sktensor/core.py:279,285).
//
// Both are expressed on the free (L, I_n, T) view of a C-order tensor -- no transpose/reshape
// copies as in the reference -- and run through one strided, batched, shared-memory tiled FP64
// GEMM:   C[b](m,n) (+)= sum_k A[b](m,k) * B[b](k,n).
#include <algorithm>

#include "common.cuh"

namespace skt {
int gram_dmma(const double *X_d, long long L, long long In, long long T, double *G_d, void *ws_d, size_t ws_bytes,
              cudaStream_t st, size_t *need_out);
int dense_impl(const double *X_d, const int64_t *shape, int ndim, const double *const *U_d, int rank, int mode,
               double *M_d, void *ws_d, size_t ws_bytes, skt_stream stream, bool rows_only);
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
    int nsplit;       // sum_batches only: grid.z CTAs share the batches; each writes the partial C + z * csplit
    long long csplit;
    long long ktotal;  // > 0: batch b only holds k with b*K + k < ktotal (ragged last batch)
};

__global__ void __launch_bounds__(GT) gemm_strided_kernel(const __grid_constant__ GemmParams p) {
    __shared__ double As[TK][TM + 1];
    __shared__ double Bs[TK][TN + 1];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    int b_lo = blockIdx.z, b_hi = blockIdx.z + 1;
    if (p.sum_batches) {
        const int per = (p.nbatch + p.nsplit - 1) / p.nsplit;
        b_lo = blockIdx.z * per;
        b_hi = min(p.nbatch, b_lo + per);
    }
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
        const int kvalid = (p.ktotal > 0) ? (int)min((long long)p.K, p.ktotal - (long long)b * p.K) : p.K;
        for (int k0 = 0; k0 < p.K; k0 += TK) {
            for (int e = tid; e < TM * TK; e += GT) {
                const int kk = a_kfast ? (e % TK) : (e / TM);
                const int mm = a_kfast ? (e / TK) : (e % TM);
                const int gm = m0 + mm, gk = k0 + kk;
                As[kk][mm] = (gm < p.M && gk < kvalid) ? A[gm * p.am + gk * p.ak] : 0.0;
            }
            for (int e = tid; e < TN * TK; e += GT) {
                const int kk = b_kfast ? (e % TK) : (e / TN);
                const int nn = b_kfast ? (e / TK) : (e % TN);
                const int gn = n0 + nn, gk = k0 + kk;
                Bs[kk][nn] = (gn < p.N && gk < kvalid) ? B[gk * p.bk + gn * p.bn] : 0.0;
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
    double *C = p.C + (p.sum_batches ? (long long)blockIdx.z * p.csplit : (long long)blockIdx.z * p.cb);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gm = m0 + ty + 16 * i, gn = n0 + tx + 16 * j;
            if (gm < p.M && gn < p.N) C[gm * p.cm + gn * p.cn] = acc[i][j];
        }
}

// M[e] = sum_s P[s][e] in split order (deterministic)
__global__ void sum_partials_kernel(const double *__restrict__ P, int nsplit, long long n, double *__restrict__ M) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int k = 0; k < nsplit; ++k) s += P[(long long)k * n + e];
        M[e] = s;
    }
}

int launch_gemm(const GemmParams &p, cudaStream_t st) {
    const int gz = p.sum_batches ? p.nsplit : p.nbatch;
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
    p.nsplit = 1; p.csplit = 0; p.ktotal = 0;
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
        // last mode with a factor-layout matrix: the rows X[(l), :] . V are exactly the GEMM rows of the dense
        // MTTKRP kernel (TMA + FP64 tensor cores), stored without weights or reduction
        if (transp && pdim <= 64) {
            const int64_t shape2[2] = {(int64_t)L, (int64_t)In};
            const double *Uv[2] = {nullptr, V_d};
            rc = dense_impl(X_d, shape2, 2, Uv, (int)pdim, 0, Y_d, nullptr, 0, stream, true);
            if (rc != SKT_EUNSUPPORTED) return rc;
        }
        // last mode:  Y (L x p) = X (L x I_n) . A^T (I_n x p)
        p.A = X_d; p.am = In; p.ak = 1; p.ab = 0;
        p.B = V_d; p.bk = a_k; p.bn = a_row; p.bb = 0;
        p.C = Y_d; p.cm = pdim; p.cn = 1; p.cb = 0;
        p.M = (int)L; p.N = (int)pdim; p.nbatch = 1;
    }
    return launch_gemm(p, (cudaStream_t)stream);
}

namespace {
// how many CTAs along z share the batches of the Gram so that the chip is filled
int gram_splits(long long In, long long nbatch) {
    const long long tiles = ((In + TM - 1) / TM) * ((In + TN - 1) / TN);
    long long s = (4LL * sm_count() + tiles - 1) / tiles;
    s = std::min<long long>(s, nbatch);
    return (int)std::max<long long>(1, std::min<long long>(s, 256));
}
void gram_view(long long L, long long In, long long T, long long &nbatch, long long &K) {
    // the contraction runs over (l, t); a batch is a slice with unit- or In*T-strided k:
    //   T > 1 : batch = l, k = t (K = T);   T == 1 (last mode): batches of 64 l, k = l within the batch
    if (T > 1) { nbatch = L; K = T; }
    else { K = std::min<long long>(64, L); nbatch = (L + K - 1) / K; }
}
}  // namespace

extern "C" size_t skt_unfold_gram_workspace(const int64_t *shape, int ndim, int mode) {
    long long L, In, T;
    if (view3(shape, ndim, mode, L, In, T) != SKT_OK) return 0;
    long long nbatch, K;
    gram_view(L, In, T, nbatch, K);
    size_t need = align_up((size_t)gram_splits(In, nbatch) * In * In * sizeof(double) + 256, 256);
    size_t need_mma = 0;
    if (gram_dmma(nullptr, L, In, T, nullptr, nullptr, 0, nullptr, &need_mma) == SKT_OK) need = std::max(need, need_mma);
    return need;
}

extern "C" int skt_unfold_gram(const double *X_d, const int64_t *shape, int ndim, int mode, double *G_d, void *ws_d,
                               size_t ws_bytes, skt_stream stream) {
    SKT_REQUIRE(X_d && G_d, SKT_EINVAL, "NULL argument");
    long long L, In, T;
    int rc = view3(shape, ndim, mode, L, In, T);
    if (rc) return rc;
    // tensor-core route (16-byte aligned rows, I_n >= 64); anything else takes the generic strided GEMM below
    rc = gram_dmma(X_d, L, In, T, G_d, ws_d, ws_bytes, (cudaStream_t)stream, nullptr);
    if (rc != SKT_EUNSUPPORTED) return rc;
    long long nbatch, K;
    gram_view(L, In, T, nbatch, K);
    const int ns = gram_splits(In, nbatch);
    const size_t need = skt_unfold_gram_workspace(shape, ndim, mode);
    SKT_REQUIRE(ns == 1 || (ws_d && ws_bytes >= need), SKT_EWORKSPACE, "workspace too small: need %zu bytes, got %zu", need,
                ws_bytes);
    // G = sum over batches of A_b B_b with A_b(i,k) = B_b(k,i) = X[.., i, ..]
    GemmParams p;
    if (T > 1) {
        p.A = X_d; p.am = T; p.ak = 1; p.ab = In * T;
        p.B = X_d; p.bk = 1; p.bn = T; p.bb = In * T;
    } else {
        // last mode: X is (L x I_n) row-major, batches of K rows; the ragged last batch is masked through ktotal
        p.A = X_d; p.am = 1; p.ak = In; p.ab = K * In;
        p.B = X_d; p.bk = In; p.bn = 1; p.bb = K * In;
    }
    p.C = (ns > 1) ? (double *)ws_d : G_d; p.cm = In; p.cn = 1; p.cb = 0;
    p.M = (int)In; p.N = (int)In; p.K = (int)K; p.nbatch = (int)nbatch; p.sum_batches = 1;
    p.nsplit = ns; p.csplit = In * In;
    p.ktotal = (T > 1) ? 0 : L;
    cudaStream_t st = (cudaStream_t)stream;
    rc = launch_gemm(p, st);
    if (rc) return rc;
    if (p.nsplit > 1) {
        const long long n = In * In;
        const int blocks = (int)std::min<long long>((n + 255) / 256, 4LL * sm_count());
        sum_partials_kernel<<<blocks, 256, 0, st>>>((const double *)ws_d, p.nsplit, n, G_d);
        SKT_LAUNCH_CHECK();
    }
    return SKT_OK;
}

// Y = X x_0 V^T with the new mode stored LAST: Y[(i_1..i_{N-1}), j] = sum_i0 X[i0, i_1..] V[i0, j].
// (dtensor._ttm_compute, dtensor.py:58-76, would give the same numbers with the new mode first; HOOI keeps
// the rotated layout between the steps of its chain, so no transpose is ever materialised.)
extern "C" int skt_ttm_dense_first_rot(const double *X_d, const int64_t *shape, int ndim, const double *V_d, int64_t pdim,
                                       double *Y_d, skt_stream stream) {
    SKT_REQUIRE(X_d && V_d && Y_d && pdim >= 1, SKT_EINVAL, "NULL argument or bad output size");
    SKT_REQUIRE(shape && ndim >= 2 && ndim <= SKT_MAX_NDIM, SKT_EINVAL, "bad shape/ndim");
    long long rest = 1;
    for (int m = 1; m < ndim; ++m) rest *= shape[m];
    SKT_REQUIRE(rest < (1LL << 31) && shape[0] < (1LL << 31), SKT_ESHAPE, "tensor view too large");
    int rc;
    if (pdim <= 64) {
        // contraction over the first mode = the M-major GEMM of the dense kernel on the (I_0, L, I_last) view,
        // rows (l, i_last) stored as they are
        int64_t shape3[3];
        const double *Uv[3] = {V_d, nullptr, nullptr};
        int nd3;
        if (ndim == 2) { shape3[0] = shape[0]; shape3[1] = shape[1]; nd3 = 2; }
        else { shape3[0] = shape[0]; shape3[1] = rest / shape[ndim - 1]; shape3[2] = shape[ndim - 1]; nd3 = 3; }
        rc = dense_impl(X_d, shape3, nd3, Uv, (int)pdim, nd3 - 1, Y_d, nullptr, 0, stream, true);
        if (rc != SKT_EUNSUPPORTED) return rc;
    }
    // generic route: Y^T (p x rest) = V^T (p x I_0) . X (I_0 x rest), written through transposed strides
    GemmParams p;
    p.sum_batches = 0; p.nsplit = 1; p.csplit = 0; p.ktotal = 0;
    p.A = V_d; p.am = 1; p.ak = pdim; p.ab = 0;
    p.B = X_d; p.bk = rest; p.bn = 1; p.bb = 0;
    p.C = Y_d; p.cm = 1; p.cn = pdim; p.cb = 0;
    p.M = (int)pdim; p.N = (int)rest; p.K = (int)shape[0]; p.nbatch = 1;
    return launch_gemm(p, (cudaStream_t)stream);
}

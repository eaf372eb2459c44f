// Block-cooperative pseudo-inverse of the Hadamard-of-Grams matrix, shared by the stand-alone
// skt_hadamard_pinv kernel and the fused small-factor update (cpals_fused.cu).
//
// Reference semantics: scipy.linalg.pinv(Y) with rtol = R * eps * sigma_max (sktensor/cp.py:144-147).
// Y is symmetric positive semi-definite (a Hadamard product of Grams).
//   fast path : unpivoted Cholesky; taken when no pivot falls below 1e-7 * max diagonal, i.e. Y is far from
//               the pinv cut-off and pinv(Y) == inv(Y) = L^-T L^-1.  One block barrier per column; L^-1 by
//               one warp per column with the running vector in registers (shuffles, no barrier).
//   general   : one-sided (Hestenes) Jacobi, one warp per column pair, round-robin ordering; singular values
//               <= R * eps * sigma_max are dropped exactly like scipy does.
#pragma once

#include <cfloat>

#include "common.cuh"

namespace skt {

struct PinvShared {
    double *W, *V;      // R x ld scratch each
    double *sig2, *dg;  // R each
    int *flags;         // 4 ints
    double *scal;       // 2 doubles
    int ld;
};

__host__ __device__ inline int pinv_ld(int R) { return R + 1 + (R & 1); }
__host__ __device__ inline size_t pinv_smem_doubles(int R) { return (size_t)2 * R * pinv_ld(R) + 2 * R + 2 + 2; }  // + flags (2 doubles)

__device__ __forceinline__ PinvShared pinv_carve(double *base, int R) {
    PinvShared s;
    s.ld = pinv_ld(R);
    s.W = base;
    s.V = s.W + (size_t)R * s.ld;
    s.sig2 = s.V + (size_t)R * s.ld;
    s.dg = s.sig2 + R;
    s.scal = s.dg + R;
    s.flags = reinterpret_cast<int *>(s.scal + 2);
    return s;
}

// Fills W with Y = hadamard_{k != mode} G_k; returns (uniformly) true when Y holds a NaN/Inf.
__device__ __forceinline__ bool pinv_load_y(const PinvShared &s, const double *__restrict__ G, int ndim, int R, int mode) {
    const int tid = threadIdx.x, nth = blockDim.x;
    if (tid == 0) s.flags[0] = 0;
    __syncthreads();
    for (int e = tid; e < R * R; e += nth) {
        const int a = e / R, b = e - a * R;
        double y = 1.0;
        for (int k = 0; k < ndim; ++k)
            if (k != mode) y *= G[(size_t)k * R * R + e];
        if (!isfinite(y)) s.flags[0] = 1;
        s.W[b * s.ld + a] = y;
    }
    __syncthreads();
    return s.flags[0] != 0;
}

// pinv of the matrix in s.W (destroyed).  `store(a, b, value)` receives every entry of P once.
// Returns the numerical rank kept (uniform over the block).  blockDim.x must be >= 32 * ceil(R/2)
// and a multiple of 32.  G/ndim/mode are only used to rebuild Y when the Cholesky path bails out.
template <class Store>
__device__ int block_pinv(const PinvShared &s, const double *__restrict__ G, int ndim, int R, int mode, Store store) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nth = blockDim.x, nw = nth >> 5;
    const int ld = s.ld;
    double *W = s.W, *V = s.V;
    constexpr unsigned FULL = 0xffffffffu;

    // ------------------------------------------------------------------ Cholesky fast path
    for (int e = tid; e < R * ld; e += nth) V[e] = W[e];   // V: working copy A(i,j) == V[j*ld+i]
    if (tid == 0) {
        double mx = 0.0;
        for (int j = 0; j < R; ++j) mx = fmax(mx, W[j * ld + j]);
        s.scal[0] = mx;
    }
    __syncthreads();
    const double maxdiag = s.scal[0];
    const double thr = 1e-7 * maxdiag;
    bool fail = !(maxdiag > 0.0);
    for (int k = 0; k < R && !fail; ++k) {
        const double d = V[k * ld + k];             // final after the previous step's update (uniform value)
        if (!(d > thr)) { fail = true; break; }
        // one reciprocal square root per column instead of a square root and two divisions (each an ~100-cycle
        // dependent sequence in FP64); dg keeps 1/L(k,k)
        const double rsd = rsqrt(d), inv_d = rsd * rsd;
        if (tid == 0) s.dg[k] = rsd;
        // trailing update with the unscaled column k; W receives column k of L
        for (int i = k + 1 + warp; i < R; i += nw) {
            const double aik = V[k * ld + i];
            if (lane == 0) W[k * ld + i] = aik * rsd;
            const double f = aik * inv_d;
            for (int j = k + 1 + lane; j <= i; j += 32) V[j * ld + i] -= f * V[k * ld + j];
        }
        __syncthreads();
    }
    if (!fail) {
        // L^-1: one warp per column j, lane r owns x_r and x_{r+32}; Linv(i,j) is stored in V[j*ld+i]
        for (int j = warp; j < R; j += nw) {
            double x0 = (lane == j) ? 1.0 : 0.0, x1 = (lane + 32 == j) ? 1.0 : 0.0;
            for (int i = lane; i < j; i += 32) V[j * ld + i] = 0.0;
            for (int i = j; i < R; ++i) {
                const double xi = __shfl_sync(FULL, (i < 32) ? x0 : x1, i & 31) * s.dg[i];
                if (lane == (i & 31)) V[j * ld + i] = xi;
                const int r0 = lane, r1 = lane + 32;
                if (r0 > i && r0 < R) x0 -= W[i * ld + r0] * xi;
                if (r1 > i && r1 < R) x1 -= W[i * ld + r1] * xi;
            }
        }
        __syncthreads();
        for (int e = tid; e < R * R; e += nth) {
            const int a = e / R, b = e - a * R;
            double sum = 0.0;
            for (int k = max(a, b); k < R; ++k) sum += V[a * ld + k] * V[b * ld + k];
            store(a, b, sum);
        }
        return R;
    }

    // ------------------------------------------------------------------ Jacobi (exact pinv truncation)
    __syncthreads();
    pinv_load_y(s, G, ndim, R, mode);               // the Cholesky attempt overwrote parts of W
    for (int e = tid; e < R * R; e += nth) {
        const int a = e / R, b = e - a * R;
        V[b * ld + a] = (a == b) ? 1.0 : 0.0;
    }
    if (tid == 0) s.flags[1] = 0;
    __syncthreads();
    const int npairs = (R + 1) / 2;
    const int m = 2 * npairs - 1;  // round-robin modulus
    if (warp == 0) {
        double mx = 0.0;
        for (int j = lane; j < R; j += 32) {
            double q = 0.0;
            for (int i = 0; i < R; ++i) q += W[j * ld + i] * W[j * ld + i];
            mx = fmax(mx, q);
        }
        mx = warp_max(mx);
        if (lane == 0) s.scal[1] = mx;
    }
    __syncthreads();
    const double eps = DBL_EPSILON;
    const double negl = s.scal[1] * (R * eps) * (R * eps) * 1e-4;  // far below the pinv cut-off

    for (int sweep = 0; sweep < 30; ++sweep) {
        for (int step = 0; step < m; ++step) {
            if (warp < npairs) {
                int p, q;
                if (warp == 0) { p = m; q = step % m; }
                else { p = (step + warp) % m; q = (step - warp + m) % m; }
                if (p > q) { const int t = p; p = q; q = t; }
                if (q < R) {
                    double *Wp = W + p * ld, *Wq = W + q * ld, *Vp = V + p * ld, *Vq = V + q * ld;
                    const double wp0 = (lane < R) ? Wp[lane] : 0.0, wq0 = (lane < R) ? Wq[lane] : 0.0;
                    const double wp1 = (lane + 32 < R) ? Wp[lane + 32] : 0.0, wq1 = (lane + 32 < R) ? Wq[lane + 32] : 0.0;
                    const double alpha = warp_sum(wp0 * wp0 + wp1 * wp1);
                    const double beta = warp_sum(wq0 * wq0 + wq1 * wq1);
                    const double gamma = warp_sum(wp0 * wq0 + wp1 * wq1);
                    const double lim = eps * sqrt(alpha * beta);
                    const bool tiny = (alpha <= negl && beta <= negl);
                    if (fabs(gamma) > lim && !tiny && alpha * beta > 0.0) {
                        const double zeta = (beta - alpha) / (2.0 * gamma);
                        const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                        const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
                        if (lane < R) {
                            Wp[lane] = c * wp0 - sn * wq0;
                            Wq[lane] = sn * wp0 + c * wq0;
                            const double vp = Vp[lane], vq = Vq[lane];
                            Vp[lane] = c * vp - sn * vq;
                            Vq[lane] = sn * vp + c * vq;
                        }
                        if (lane + 32 < R) {
                            Wp[lane + 32] = c * wp1 - sn * wq1;
                            Wq[lane + 32] = sn * wp1 + c * wq1;
                            const double vp = Vp[lane + 32], vq = Vq[lane + 32];
                            Vp[lane + 32] = c * vp - sn * vq;
                            Vq[lane + 32] = sn * vp + c * vq;
                        }
                        if (lane == 0) s.flags[1] = 1;
                    }
                }
            }
            __syncthreads();
        }
        const int any = s.flags[1];
        __syncthreads();
        if (tid == 0) s.flags[1] = 0;
        __syncthreads();
        if (!any) break;
    }
    for (int j = warp; j < R; j += nw) {
        double q = 0.0;
        for (int i = lane; i < R; i += 32) q += W[j * ld + i] * W[j * ld + i];
        q = warp_sum(q);
        if (lane == 0) s.sig2[j] = q;
    }
    __syncthreads();
    if (tid == 0) {
        double mx = 0.0;
        for (int j = 0; j < R; ++j) mx = fmax(mx, s.sig2[j]);
        const double cut = (double)R * eps * sqrt(mx);
        int kept = 0;
        for (int j = 0; j < R; ++j) {
            if (sqrt(s.sig2[j]) > cut) ++kept;
            else s.sig2[j] = 0.0;  // dropped
        }
        s.flags[2] = kept;
    }
    __syncthreads();
    for (int e = tid; e < R * R; e += nth) {
        const int a = e / R, b = e - a * R;
        double q = 0.0;
        for (int j = 0; j < R; ++j)
            if (s.sig2[j] > 0.0) q += V[j * ld + a] * W[j * ld + b] / s.sig2[j];
        store(a, b, q);
    }
    return s.flags[2];
}

}  // namespace skt

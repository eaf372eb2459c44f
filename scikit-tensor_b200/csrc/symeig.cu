// Leading eigenvectors of a small symmetric matrix in ONE kernel on ONE thread-block cluster (sm_100a).
//
// Replaces the LAPACK call inside sktensor.core.nvecs (reference: sktensor/core.py:275-294, `eigh(Y, eigvals=(N - rank,
// N - 1))`, columns reversed to descending order, then flipsign, core.py:297-306) -- the step that follows every TTM
// chain of tucker.hooi (tucker.py:103-104) and the default initialisation of cp.als (cp.py:200-202).  A library
// eigensolver spends ~4 ms on the 384 x 384 problem of BASELINE config 4 because its tridiagonalisation is a chain of
// ~n dependent small launches; here the matrix lives in the DISTRIBUTED SHARED MEMORY of one cluster (column-cyclic over
// up to 16 CTAs) and the n Householder steps are separated by cluster barriers (~0.2 us) instead of launches:
//
//   1. Householder tridiagonalisation A = Q T Q^T.  Per step: the owner of column k forms the reflector and broadcasts it
//      to every CTA's shared memory (DSMEM stores); every CTA applies the PREVIOUS step's rank-2 update to its columns and,
//      in the same pass, forms its entries of p = tau A v; p is all-gathered through DSMEM; w = p - (tau/2 p.v) v is
//      formed redundantly.  Two cluster barriers per step, one pass over the trailing matrix per step.
//   2. The k largest eigenvalues of T by multisection on Sturm counts (eigenvalues spread over the CTAs, up to 256 shifts
//      per eigenvalue and round).
//   3. Their eigenvectors by inverse iteration on the pivoted LU of T - lambda I (dlagtf / dlagts arithmetic, one thread
//      per eigenvalue), with modified Gram-Schmidt inside groups of close eigenvalues after every solve (dstein's recipe,
//      all vectors advanced together).
//   4. Back-transformation V = Q Z (reflectors streamed from L2 through a cp.async ring), descending order, flipsign.
//
// tools/symeig_model.py is a NumPy model of exactly these recurrences; tests/test_symeig_model.py pins it to LAPACK.
#include <algorithm>
#include <cfloat>
#include <cstdlib>

#include "common.cuh"

namespace skt {
namespace {

constexpr int ET = 512;          // threads per CTA
constexpr int EW = ET / 32;
constexpr int RB = 8;            // reflectors per chunk of the back-transformation ring
constexpr int EIG_MAX_N = 2048;
constexpr int LW = 32;            // largest group that takes the symmetric (Loewdin) orthonormalisation
constexpr int DS = 8;             // threads per dot product in the back-transformation
constexpr int ZCAP_MAX = 64;      // vectors per back-transformation pass (sizes the small WY scratch)

struct EigParams {
    const double *A;
    int n, k, flip;
    double *W;        // [k] descending (nullable)
    double *U;        // [n][k] row-major
    int *info;        // nullable
    double *Vg;       // [n][ldv] reflectors (column kk at Vg + kk * ldv)
    double *Zg;       // [k][n]
    double *lamg;     // [k]
    double *Ag;       // [C][ncl * n] column blocks when they do not fit shared memory (else null)
    double *Fg;       // [C][kc * fslot] factor arrays when they do not fit shared memory (else null)
    int C, ncl, kc, L, ldv, z_smem, iters, zcap, cholqr;
    unsigned long long *prof;   // optional phase timestamps (globaltimer ns) written by CTA 0, SKT_EIG_PROF=1
};

// shared-memory layout in doubles (host and device agree through this function)
struct EigLayout {
    int vb, wb, pb, dd, ee, tt, lam, grp, tgrp, red, nbs, wy, big;
    size_t big_doubles;      // capacity of the phase-overlaid region
    size_t total_bytes;
};
__host__ __device__ inline int fslot_doubles(int n) { return 5 * n + (n + 7) / 8 + 1; }
__host__ __device__ inline EigLayout eig_layout(int n, int k, size_t big_doubles) {
    EigLayout L;
    int o = 0;
    L.vb = o; o += 2 * n;
    L.wb = o; o += 2 * n;
    L.pb = o; o += n;
    L.dd = o; o += n;
    L.ee = o; o += n;
    L.tt = o; o += n;
    L.lam = o; o += k;
    L.grp = o; o += (k + 1) / 2 + 1;        // ints
    L.tgrp = o; o += (k + 1) / 2 + 1;       // ints
    L.red = o; o += 2 * EW + 2;
    L.nbs = o; o += (ET + 1) / 2 + 1;       // ints (one per multisection slot)
    L.wy = o; o += 2 * RB * RB + 2 * RB * ZCAP_MAX + ET;   // compact-WY scratch: G, T, V^T z, y, partial dot products
    o = (o + 1) & ~1;                       // 16-byte alignment of the big region
    L.big = o;
    L.big_doubles = big_doubles;
    L.total_bytes = ((size_t)o + big_doubles) * sizeof(double);
    return L;
}

__device__ __forceinline__ double hash_uniform(unsigned long long i, unsigned long long j) {
    unsigned long long h = i * 0x9E3779B97F4A7C15ull + j * 0xC2B2AE3D27D4EB4Full + 0x165667B19E3779F9ull;
    h ^= h >> 29;
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 32;
    return (double)(h >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}

// 1 / x from the hardware approximation and two Newton steps (full precision, not correctly rounded): the Sturm and
// LU recurrences are chains of dependent divisions, and an IEEE division costs ~200 cycles of latency per step
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;\n" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

__global__ void __launch_bounds__(ET, 1) symeig_cluster_kernel(const __grid_constant__ EigParams p) {
    extern __shared__ __align__(16) double esm[];
    const int n = p.n, k = p.k, C = p.C, ncl = p.ncl, kc = p.kc;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const EigLayout Lo = eig_layout(n, k, 0);
    double *vb = esm + Lo.vb, *wb = esm + Lo.wb, *pb = esm + Lo.pb, *dd = esm + Lo.dd, *ee = esm + Lo.ee, *tt = esm + Lo.tt;
    double *lam = esm + Lo.lam, *red = esm + Lo.red, *big = esm + Lo.big;
    int *grp = reinterpret_cast<int *>(esm + Lo.grp);
    int *tgrp = reinterpret_cast<int *>(esm + Lo.tgrp);
    int *nbs = reinterpret_cast<int *>(esm + Lo.nbs);
    unsigned crank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(crank));
    const int c = (int)crank;

    auto stamp = [&](int slot) {
        if (p.prof && c == 0 && tid == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
            p.prof[slot] = t;
        }
    };
    auto cluster_arrive = [&]() { asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory"); };
    auto cluster_wait = [&]() { asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory"); };
    auto cluster_sync = [&]() {
        asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    };
    auto remote = [&](double *ptr, int rank) -> double * {
        uint64_t in = (uint64_t)(uintptr_t)ptr, out;
        asm volatile("mapa.u64 %0, %1, %2;\n" : "=l"(out) : "l"(in), "r"((unsigned)rank));
        return (double *)(uintptr_t)out;
    };
    // block-wide reductions with a fixed order (every thread returns the same value)
    auto block_sum = [&](double v) -> double {
        v = warp_sum(v);
        __syncthreads();
        if (lane == 0) red[warp] = v;
        __syncthreads();
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < EW; ++w) s += red[w];
        return s;
    };
    auto block_max = [&](double v) -> double {
        v = warp_max(v);
        __syncthreads();
        if (lane == 0) red[warp] = v;
        __syncthreads();
        double s = red[0];
#pragma unroll
        for (int w = 1; w < EW; ++w) s = fmax(s, red[w]);
        return s;
    };

    // ================================================================== 1. tridiagonalisation
    double *Ablk = p.Ag ? p.Ag + (size_t)c * ncl * n : big;
    for (int t = 0; t < ncl; ++t) {
        const int j = c + C * t;
        if (j >= n) break;
        for (int i = tid; i < n; i += ET) Ablk[(size_t)t * n + i] = p.A[(size_t)j * n + i];   // row j == column j
    }
    __syncthreads();
    cluster_sync();                       // every CTA is resident before the first remote store

    stamp(0);
    // wb doubles as the second p buffer: the rank-2 update of step kk-1, A -= v w^T + w v^T with w = p + al v, is applied
    // on the fly from (v, p, al) during step kk's pass, so w is never materialised and no CTA-wide reduction is needed
    bool has_pend = false;
    double alo = 0.0;                              // al of the pending update (identical in every thread)
    for (int kk = 0; kk + 1 < n; ++kk) {
        const int cur = kk & 1, prev = cur ^ 1;
        double *vc = vb + cur * n, *vo = vb + prev * n;
        double *pc = wb + cur * n, *po = wb + prev * n;
        if (c == kk % C) {
            double *col = Ablk + (size_t)(kk / C) * n;
            // every thread owns one row of the column: update it, form the reflector, broadcast it
            const double vk = has_pend ? vo[kk] : 0.0, wk = has_pend ? po[kk] + alo * vo[kk] : 0.0;
            double part = 0.0;
            for (int i = kk + tid; i < n; i += ET) {
                double a = col[i];
                if (has_pend) {
                    a -= vo[i] * wk + (po[i] + alo * vo[i]) * vk;
                    col[i] = a;
                }
                if (i >= kk + 2) part += a * a;
            }
            const double xnorm2 = block_sum(part);        // (its barriers also publish col[])
            const double alpha = col[kk + 1], dk = col[kk];
            double tau, ek, scal;
            if (xnorm2 == 0.0) {
                tau = 0.0; ek = alpha; scal = 0.0;
            } else {
                const double beta = -copysign(sqrt(alpha * alpha + xnorm2), alpha);
                tau = (beta - alpha) * fast_rcp(beta);
                scal = fast_rcp(alpha - beta);
                ek = beta;
            }
            __syncthreads();                              // alpha / dk were read by everyone before col[kk + 1] changes
            for (int i = kk + 1 + tid; i < n; i += ET) {
                const double v = (i == kk + 1) ? 1.0 : col[i] * scal;
                col[i] = v;                                // the reflector replaces the eliminated column (dumped to Vg below)
                for (int r = 0; r < C; ++r) remote(vc, r)[i] = v;
            }
            if (tid < C) {
                remote(dd, tid)[kk] = dk;
                remote(ee, tid)[kk] = ek;
                remote(tt, tid)[kk] = tau;
            }
            cluster_arrive();                             // v is on its way; the local update below overlaps the barrier
        }
        // previous step's rank-2 update on the columns this CTA owns.  It needs nothing from the current step, so the
        // other CTAs do it WHILE the owner forms the reflector, and the owner between its arrive and its wait.
        const int t0 = max((kk + 1 - c + C - 1) / C, 0);  // first local column with j >= kk + 1  (j = c + C t)
        if (has_pend) {
            for (int t = t0 + warp; t < ncl; t += EW) {
                const int j = c + C * t;
                if (j >= n) break;
                double *col = Ablk + (size_t)t * n;
                const double vj = vo[j], wj = po[j] + alo * vo[j];
                if (p.Ag) {                                         // columns in L2: keep four loads in flight per lane
#pragma unroll 4
                    for (int i = kk + 1 + lane; i < n; i += 32) {
                        const double voi = vo[i];
                        col[i] -= voi * wj + (po[i] + alo * voi) * vj;
                    }
                } else {
                    for (int i = kk + 1 + lane; i < n; i += 32) {
                        const double voi = vo[i];
                        col[i] -= voi * wj + (po[i] + alo * voi) * vj;
                    }
                }
            }
        }
        if (c != kk % C) cluster_arrive();
        cluster_wait();
        const double tau = tt[kk];
        // this step's p = tau A v on the same columns (same warp, same lanes: no barrier in between)
        for (int t = t0 + warp; t < ncl; t += EW) {
            const int j = c + C * t;
            if (j >= n) break;
            const double *col = Ablk + (size_t)t * n;
            double acc = 0.0;
            if (p.Ag) {                                             // columns in L2: four loads in flight per lane
                double acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
                int i = kk + 1 + lane;
                for (; i + 96 < n; i += 128) {
                    acc += col[i] * vc[i];
                    acc1 += col[i + 32] * vc[i + 32];
                    acc2 += col[i + 64] * vc[i + 64];
                    acc3 += col[i + 96] * vc[i + 96];
                }
                for (; i < n; i += 32) acc += col[i] * vc[i];
                acc = (acc + acc1) + (acc2 + acc3);
            } else {
                for (int i = kk + 1 + lane; i < n; i += 32) acc += col[i] * vc[i];
            }
            acc = warp_sum(acc) * tau;
            if (lane < C) remote(pc, lane)[j] = acc;
        }
        cluster_sync();
        if (tau != 0.0) {
            double part = 0.0;                      // al = -tau/2 p.v, by every warp for itself (same order everywhere)
            for (int i = kk + 1 + lane; i < n; i += 32) part += pc[i] * vc[i];
            alo = -0.5 * tau * warp_sum(part);
            has_pend = true;
        } else {
            has_pend = false;
        }
    }
    // the reflectors, kept in place in their owners' columns, go to global memory for the back-transformation
    for (int t = warp; t < ncl; t += EW) {
        const int j = c + C * t;
        if (j >= n - 1) break;
        for (int i = j + 1 + lane; i < n; i += 32) p.Vg[(size_t)j * p.ldv + i] = Ablk[(size_t)t * n + i];
    }
    // last diagonal entry (the final step's reflector is always the identity, so nothing is pending)
    if (c == (n - 1) % C && tid < C) {
        const double dl = Ablk[(size_t)((n - 1) / C) * n + (n - 1)];
        remote(dd, tid)[n - 1] = dl;
        if (n >= 1) { remote(ee, tid)[n - 1] = 0.0; remote(tt, tid)[n - 1] = 0.0; }
    }
    cluster_sync();
    stamp(1);

    // ================================================================== 2. eigenvalues (multisection on Sturm counts)
    const double EPS = DBL_EPSILON, SAFMIN = DBL_MIN;
    double gl, gu, pivmin, bnorm;
    {
        double lo = DBL_MAX, hi = -DBL_MAX, em = 0.0;
        for (int i = tid; i < n; i += ET) {
            const double el = i > 0 ? fabs(ee[i - 1]) : 0.0, er = i + 1 < n ? fabs(ee[i]) : 0.0;
            lo = fmin(lo, dd[i] - el - er);
            hi = fmax(hi, dd[i] + el + er);
            em = fmax(em, er * er);
        }
        for (int i = tid; i + 1 < n; i += ET) pb[i] = ee[i] * ee[i];       // e^2 for the Sturm recurrence
        gu = block_max(hi);
        gl = -block_max(-lo);
        em = block_max(em);
        pivmin = SAFMIN * fmax(1.0, em);
        bnorm = fmax(fabs(gl), fabs(gu));
        gl -= 2.1 * bnorm * EPS * n + 2.1 * pivmin;
        gu += 2.1 * bnorm * EPS * n + 2.1 * pivmin;
    }
    {
        const int L = p.L;
        const int s = tid / L, l = tid - s * L;
        const int q = c + C * s;
        const bool active = (s < kc) && (q < k);
        const int idx = n - k + q;
        double lo = gl, hi = gu;
        const int max_rounds = 70;
        for (int round = 0; round < max_rounds; ++round) {
            const bool conv = (hi - lo) <= 2.0 * EPS * fmax(fabs(lo), fabs(hi)) + 2.0 * pivmin;
            if (__syncthreads_and((conv || !active) ? 1 : 0)) break;
            if (tid < ET / L) nbs[tid] = 0;
            __syncthreads();
            const double w = hi - lo;
            if (active && !conv) {
                const double x = lo + w * ((l + 1.0) / (L + 1.0));
                double qv = dd[0] - x;
                if (fabs(qv) < pivmin) qv = -pivmin;
                int cnt = qv < 0.0;
                for (int i = 1; i < n; ++i) {
                    qv = (dd[i] - x) - pb[i - 1] * fast_rcp(qv);          // pb holds e^2 (filled above)
                    if (fabs(qv) < pivmin) qv = -pivmin;
                    cnt += qv < 0.0;
                }
                if (cnt <= idx) atomicAdd(&nbs[s], 1);
            }
            __syncthreads();
            if (active && !conv) {
                const int nb = nbs[s];
                const double nlo = nb > 0 ? lo + w * (nb / (L + 1.0)) : lo;
                const double nhi = nb < L ? lo + w * ((nb + 1.0) / (L + 1.0)) : hi;
                lo = fmax(lo, nlo);
                hi = fmin(hi, nhi);
            }
        }
        if (active && l == 0) {
            const double lq = 0.5 * (lo + hi);
            p.lamg[q] = lq;
            if (p.W) p.W[k - 1 - q] = lq;
        }
    }
    __threadfence();
    cluster_sync();
    stamp(2);
    for (int j = tid; j < k; j += ET) lam[j] = __ldcg(p.lamg + j);
    __syncthreads();
    const double onenrm = fmax(bnorm, SAFMIN);
    if (tid == 0) {
        const double ortol = 1e-3 * onenrm;
        grp[0] = 0;
        tgrp[0] = 0;
        for (int j = 1; j < k; ++j) {
            const double sep = 10.0 * EPS * fabs(lam[j]);
            if (lam[j] - lam[j - 1] < sep) lam[j] = lam[j - 1] + sep;
            grp[j] = (lam[j] - lam[j - 1] < ortol) ? grp[j - 1] : j;
            tgrp[j] = (lam[j] - lam[j - 1] < 1e-8 * onenrm) ? tgrp[j - 1] : j;     // numerically coincident
        }
    }
    __syncthreads();

    // ================================================================== 3. eigenvectors of T (inverse iteration)
    const int fslot = fslot_doubles(n);
    double *Freg = p.Fg ? p.Fg + (size_t)c * kc * fslot : big;
    double *Zs = p.z_smem ? big + (p.Fg ? 0 : (size_t)kc * fslot) : p.Zg;     // CTA 0's orthogonalisation copy
    const int myk = (k > c) ? (k - c + C - 1) / C : 0;                        // eigenvalues owned by this CTA: q = c + C s
    // One WARP per eigenvalue.  The recurrences are sequential in the row index, so every lane runs the same chain on
    // operands that the lanes fetched 32 rows at a time (coalesced shared-memory loads) and hand round by shuffles; lane t
    // keeps the outputs of row t and the chunk is stored coalesced.  (One thread per eigenvalue walking shared memory
    // took ~300 cycles per row.)
    auto alast_of = [&](int s) -> double { return fabs(Freg[(size_t)s * fslot + 3 * n + (n - 1)]); };   // parked |U_nn|, see below
    for (int s = warp; s < myk; s += EW) {
        // ---- dlagtf: T - lambda I = P L U
        const int q = c + C * s;
        double *ra = Freg + (size_t)s * fslot, *bb = ra + n, *cc = bb + n, *d2 = cc + n, *z = d2 + n;
        unsigned char *sw = reinterpret_cast<unsigned char *>(d2 + 2 * n);
        const double lq = lam[q];
        double ai = dd[0] - lq, bi = n > 1 ? ee[0] : 0.0, amax = 0.0;
        double scale1 = fabs(ai) + fabs(bi);
        for (int i0 = 0; i0 + 1 < n; i0 += 32) {
            const int il = i0 + lane;
            const double l_ci = (il + 1 < n) ? ee[il] : 0.0;
            const double l_a1 = (il + 1 < n) ? dd[il + 1] - lq : 0.0;
            const double l_b1 = (il + 2 < n) ? ee[il + 1] : 0.0;
            double k_ui = 0.0, k_bo = 0.0, k_m = 0.0, k_di = 0.0;
            unsigned char k_sw = 0;
            const int cnt = min(32, n - 1 - i0);
            for (int t = 0; t < cnt; ++t) {
                const double ci = shfl_d(l_ci, t), a1 = shfl_d(l_a1, t), b1 = shfl_d(l_b1, t);
                const double scale2 = fabs(ci) + fabs(a1) + fabs(b1);
                const bool noswap = (ci == 0.0) || (ai != 0.0 && fabs(ci) * scale1 <= fabs(ai) * scale2);
                double m, ui, bo, di, an, bn;
                if (noswap) {
                    m = (ci == 0.0) ? 0.0 : ci * fast_rcp(ai);
                    ui = ai; bo = bi; di = 0.0;
                    an = a1 - m * bi; bn = b1;
                } else {
                    m = ai * fast_rcp(ci);
                    ui = ci; bo = a1; di = b1;
                    an = bi - m * a1; bn = -m * b1;
                }
                if (lane == t) { k_ui = ui; k_bo = bo; k_m = m; k_di = di; k_sw = noswap ? 0 : 1; }
                amax = fmax(amax, fmax(fabs(ui), fmax(fabs(bo), fabs(di))));
                ai = an; bi = bn; scale1 = scale2;
            }
            if (il + 1 < n) { ra[il] = k_ui; bb[il] = k_bo; cc[il] = k_m; d2[il] = k_di; sw[il] = k_sw; }
        }
        amax = fmax(amax, fabs(ai));
        const double tol = (amax == 0.0) ? EPS : fmax(amax * EPS, SAFMIN / EPS);
        __syncwarp();
        if (lane == 0) { ra[n - 1] = ai; d2[n - 1] = ai; }     // d2 has n slots, U uses n - 2: |U_nn| parked for the scaling
        __syncwarp();
        // pivots below tol are replaced by +-tol (dlagts, job = -1) and inverted once; start vector
        for (int i = lane; i < n; i += 32) {
            const double a = ra[i];
            ra[i] = 1.0 / (fabs(a) < tol ? (a < 0.0 ? -tol : tol) : a);
            z[i] = hash_uniform((unsigned long long)i, (unsigned long long)q);
        }
    }
    __syncthreads();
    stamp(3);
    for (int it = 0; it < p.iters; ++it) {
        for (int s = warp; s < myk; s += EW) {
            const int q = c + C * s;
            double *ra = Freg + (size_t)s * fslot, *bb = ra + n, *cc = bb + n, *d2 = cc + n, *z = d2 + n;
            const unsigned char *sw = reinterpret_cast<const unsigned char *>(d2 + 2 * n);
            double nrm1 = 0.0;
            for (int i = lane; i < n; i += 32) nrm1 += fabs(z[i]);
            nrm1 = warp_sum(nrm1);
            const double scl = n * onenrm * fmax(EPS, alast_of(s)) / (nrm1 == 0.0 ? 1.0 : nrm1);
            // forward: apply P L   (y_{i-1} is final once row i has been seen)
            double yp = z[0] * scl;
            for (int i0 = 1; i0 < n; i0 += 32) {
                const int il = i0 + lane;
                const double l_z = il < n ? z[il] * scl : 0.0;
                const double l_m = il < n ? cc[il - 1] : 0.0;
                const int l_s = il < n ? sw[il - 1] : 0;
                double keep = 0.0;
                const int cnt = min(32, n - i0);
                for (int t = 0; t < cnt; ++t) {
                    const double yi = shfl_d(l_z, t), m = shfl_d(l_m, t);
                    const int sflag = __shfl_sync(0xffffffffu, l_s, t);
                    double out;
                    if (sflag) { out = yi; yp = yp - m * yi; } else { out = yp; yp = yi - m * yp; }
                    if (lane == t) keep = out;
                }
                __syncwarp();
                if (il < n) z[il - 1] = keep;
            }
            __syncwarp();
            if (lane == 0) z[n - 1] = yp;
            __syncwarp();
            // backward: U x = y
            double y1 = 0.0, y2 = 0.0, zmax = 0.0;
            for (int ih = n - 1; ih >= 0; ih -= 32) {
                const int il = ih - lane;
                const double l_z = il >= 0 ? z[il] : 0.0;
                const double l_b = (il >= 0 && il + 1 < n) ? bb[il] : 0.0;
                const double l_d = (il >= 0 && il + 2 < n) ? d2[il] : 0.0;
                const double l_r = il >= 0 ? ra[il] : 0.0;
                double keep = 0.0;
                const int cnt = min(32, ih + 1);
                for (int t = 0; t < cnt; ++t) {
                    double v = shfl_d(l_z, t) - shfl_d(l_b, t) * y1 - shfl_d(l_d, t) * y2;
                    v *= shfl_d(l_r, t);
                    if (lane == t) keep = v;
                    y2 = y1; y1 = v;
                    zmax = fmax(zmax, fabs(v));
                }
                if (il >= 0) z[il] = keep;
            }
            __syncwarp();
            const double sc = (zmax > 0.0 && zmax < DBL_MAX) ? 1.0 / zmax : 1.0;
            double *zg = p.Zg + (size_t)q * n;
            for (int i = lane; i < n; i += 32) zg[i] = z[i] * sc;
        }
        if (it + 1 == p.iters) stamp(10);
        __threadfence();
        cluster_sync();
        if (it + 1 == p.iters) stamp(11);
        if (c == 0) {
            // ---- orthonormalisation inside groups of close eigenvalues (dstein's ortol on the last pass; between the
            // solves only numerically coincident eigenvalues need it).  Vectors that come out of the solve nearly
            // orthonormal -- the usual case -- get the symmetric third-order correction Z <- Z (I - E/2 + 3/8 E^2),
            // E = Z^T Z - I: every dot product at once, no chain of dependent projections.  Anything else (tight
            // clusters, large groups) goes through modified Gram-Schmidt in ascending order.
            const bool last = (it + 1 == p.iters);
            const int *gsel = last ? grp : tgrp;
            if (p.z_smem) {
                for (size_t e = tid; e < (size_t)k * n; e += ET) Zs[e] = __ldcg(p.Zg + e);
                __syncthreads();
            }
            auto normalise = [&](int j0, int j1, bool store) {
                for (int j = j0 + warp; j < j1; j += EW) {
                    double *zj = Zs + (size_t)j * n;
                    double nn = 0.0;
                    for (int r = lane; r < n; r += 32) { const double v = p.z_smem ? zj[r] : __ldcg(zj + r); nn += v * v; }
                    nn = warp_sum(nn);
                    const double f = nn > 0.0 ? rsqrt(nn) : 0.0;
                    for (int r = lane; r < n; r += 32) {
                        const double v = (p.z_smem ? zj[r] : __ldcg(zj + r)) * f;
                        zj[r] = v;
                        if (store && p.z_smem) p.Zg[(size_t)j * n + r] = v;
                    }
                }
            };
            normalise(0, k, false);
            __syncthreads();
            for (int g0 = 0, g1 = 0; g0 < k; g0 = g1) {
                g1 = g0 + 1;
                while (g1 < k && gsel[g1] == gsel[g0]) ++g1;
                const int mg = g1 - g0;
                if (mg == 1) continue;
                bool done = false;
                if (p.cholqr && last && mg <= LW) {
                    double *Es = Freg, *Ms = Es + LW * LW;   // [LW][LW] each, over the LU factors (dead after the last solve)
                    int *cflag = nbs;
                    if (tid == 0) cflag[0] = 0;
                    __syncthreads();
                    for (int task = tid; task < mg * mg; task += ET) {
                        const int a_ = task / mg, b_ = task - a_ * mg;
                        if (a_ > b_) continue;
                        const double *za = Zs + (size_t)(g0 + a_) * n, *zb = Zs + (size_t)(g0 + b_) * n;
                        double d = 0.0;
                        int r = lane % n;                    // staggered start: the threads of a warp hit different banks
                        for (int cnt = 0; cnt < n; ++cnt) {
                            d += za[r] * zb[r];
                            r = (r + 1 == n) ? 0 : r + 1;
                        }
                        d -= (a_ == b_) ? 1.0 : 0.0;
                        Es[a_ * LW + b_] = d;
                        Es[b_ * LW + a_] = d;
                        if (!(fabs(d) < 1e-4)) cflag[0] = 1;
                    }
                    __syncthreads();
                    if (cflag[0] == 0) {
                        for (int task = tid; task < mg * mg; task += ET) {
                            const int a_ = task / mg, b_ = task - a_ * mg;
                            double e2 = 0.0;
                            for (int q_ = 0; q_ < mg; ++q_) e2 += Es[a_ * LW + q_] * Es[q_ * LW + b_];
                            Ms[a_ * LW + b_] = (a_ == b_ ? 1.0 : 0.0) - 0.5 * Es[a_ * LW + b_] + 0.375 * e2;
                        }
                        __syncthreads();
                        for (int r = tid; r < n; r += ET) {  // one row of Z_g per thread: z_row <- z_row M
                            double zr[LW];
#pragma unroll
                            for (int i = 0; i < LW; ++i) zr[i] = i < mg ? Zs[(size_t)(g0 + i) * n + r] : 0.0;
                            for (int j = 0; j < mg; ++j) {
                                double v = 0.0;
#pragma unroll
                                for (int i = 0; i < LW; ++i) v += zr[i] * Ms[i * LW + j];
                                Zs[(size_t)(g0 + j) * n + r] = v;
                            }
                        }
                        done = true;
                    }
                    __syncthreads();
                }
                if (p.prof && tid == 0 && last) p.prof[done ? 14 : 15] += (unsigned long long)mg;
                if (done) continue;
                for (int i = g0; i + 1 < g1; ++i) {
                    const double *zi = Zs + (size_t)i * n;
                    double nn = 0.0;
                    for (int r = lane; r < n; r += 32) { const double v = p.z_smem ? zi[r] : __ldcg(zi + r); nn += v * v; }
                    nn = warp_sum(nn);
                    for (int j = i + 1 + warp; j < g1; j += EW) {
                        double *zj = Zs + (size_t)j * n;
                        double dot = 0.0;
                        for (int r = lane; r < n; r += 32)
                            dot += p.z_smem ? zi[r] * zj[r] : __ldcg(zi + r) * __ldcg(zj + r);
                        dot = warp_sum(dot);
                        const double f = nn > 0.0 ? dot / nn : 0.0;
                        for (int r = lane; r < n; r += 32)
                            zj[r] = p.z_smem ? zj[r] - f * zi[r] : __ldcg(zj + r) - f * __ldcg(zi + r);
                    }
                    __syncthreads();
                }
            }
            __syncthreads();
            normalise(0, k, true);
        }
        if (it + 1 == p.iters) stamp(12);
        __threadfence();
        cluster_sync();
        if (it + 1 == p.iters) stamp(13);
        if (it + 1 < p.iters) {
            for (int s = warp; s < myk; s += EW) {
                const int q = c + C * s;
                double *z = Freg + (size_t)s * fslot + 4 * n;
                const double *zg = p.Zg + (size_t)q * n;
                for (int i = lane; i < n; i += 32) z[i] = __ldcg(zg + i);
            }
            __syncthreads();
        }
        stamp(4 + it);
    }

    // ================================================================== 4. back-transformation, order, sign
    // X = H_0 H_1 ... H_{n-3} Z, RB reflectors at a time in compact WY form: H_lo ... H_hi = I - V T V^T (dlarft, forward
    // column-wise).  Per chunk: all dot products V^T V and V^T z in parallel over the warps, T and y = T V^T z by a few
    // threads, then one rank-RB update of every vector -- three block barriers per RB reflectors instead of a dependent
    // dot / update chain per reflector.  Vectors are processed `zcap` at a time when they do not all fit.
    {
        const int ldv = p.ldv;
        const int nref = n - 2;                          // reflectors 0 .. n-3 (the last step's is the identity)
        const int nchunk = nref > 0 ? (nref + RB - 1) / RB : 0;
        double *ring = big;                              // [2][RB][ldv]
        double *zv = big + (size_t)2 * RB * ldv;         // [zcap][n]
        double *Gs = esm + Lo.wy;                        // [RB][RB] V^T V
        double *Ts = Gs + RB * RB;                       // [RB][RB]
        double *ws = Ts + RB * RB;                       // [zcap][RB] V^T z
        double *ys = ws + RB * ZCAP_MAX;                 // [zcap][nb] y = T V^T z
        double *parts = ys + RB * ZCAP_MAX;              // [ET] partial dot products
        const int zcap = p.zcap;
        bool bad = false;
        for (int s0 = 0; s0 < myk; s0 += zcap) {
            const int nz = min(zcap, myk - s0);
            __syncthreads();
            for (int s = 0; s < nz; ++s) {
                const int q = c + C * (s0 + s);
                for (int i = tid; i < n; i += ET) zv[(size_t)s * n + i] = __ldcg(p.Zg + (size_t)q * n + i);
            }
            auto issue_chunk = [&](int ch) {             // chunk ch holds reflectors lo .. lo+nb-1 (slot j = kk - lo)
                double *dst = ring + (size_t)(ch & 1) * RB * ldv;
                const int hi = nref - 1 - ch * RB, lo = max(hi - RB + 1, 0);
                const int per = ldv / 2;                 // 16-byte pieces per column
                for (int e = tid; e < (hi - lo + 1) * per; e += ET) {
                    const int j = e / per, piece = e - j * per;
                    cp_async16(dst + (size_t)j * ldv + 2 * piece, p.Vg + (size_t)(lo + j) * ldv + 2 * piece, 16);
                }
                cp_async_commit();
            };
            if (nchunk > 0) issue_chunk(0);
            for (int ch = 0; ch < nchunk; ++ch) {
                if (ch + 1 < nchunk) issue_chunk(ch + 1); else cp_async_commit();
                cp_async_wait<1>();
                __syncthreads();
                const double *V = ring + (size_t)(ch & 1) * RB * ldv;
                const int hi = nref - 1 - ch * RB, lo = max(hi - RB + 1, 0), nb = hi - lo + 1;
                // ---- A. dot products: pairs (a < b) of reflectors, and (reflector, vector); rows below a reflector's
                // start hold stale memory and are never read (v_j lives on rows >= lo + j + 1, v_j[lo + j + 1] = 1)
                const int npair = nb * (nb - 1) / 2, ntask = npair + nb * nz;
                // every dot product is split over DS threads (a warp reduction costs ~220 ns on this part, a shared
                // memory round trip a fraction of it); 512 / DS tasks per round, partial sums added in a fixed order
                for (int t0_ = 0; t0_ < ntask; t0_ += ET / DS) {
                    const int task = t0_ + tid / DS, sub = tid % DS;
                    double d = 0.0;
                    if (task < ntask) {
                        if (task < npair) {
                            int a_ = 0, rem = task;
                            while (rem >= nb - 1 - a_) { rem -= nb - 1 - a_; ++a_; }
                            const int b_ = a_ + 1 + rem;
                            const double *va = V + (size_t)a_ * ldv, *vb_ = V + (size_t)b_ * ldv;
                            for (int i = lo + b_ + 1 + sub; i < n; i += DS) d += va[i] * vb_[i];
                        } else {
                            const int t2 = task - npair, s = t2 / nb, j = t2 - s * nb;
                            const double *vj = V + (size_t)j * ldv, *z = zv + (size_t)s * n;
                            for (int i = lo + j + 1 + sub; i < n; i += DS) d += vj[i] * z[i];
                        }
                    }
                    parts[tid] = d;
                    __syncthreads();
                    if (tid < ET / DS && t0_ + tid < ntask) {
                        double sum = 0.0;
#pragma unroll
                        for (int q_ = 0; q_ < DS; ++q_) sum += parts[tid * DS + q_];
                        const int tk = t0_ + tid;
                        if (tk < npair) {
                            int a_ = 0, rem = tk;
                            while (rem >= nb - 1 - a_) { rem -= nb - 1 - a_; ++a_; }
                            Gs[a_ * RB + a_ + 1 + rem] = sum;
                        } else {
                            const int t2 = tk - npair, s = t2 / nb, j = t2 - s * nb;
                            ws[s * RB + j] = sum;
                        }
                    }
                    __syncthreads();
                }
                // ---- B. coefficients of the rank-nb update without forming T: applying H_hi first, then H_hi-1, ...,
                // s_j = tau_j v_j^T (z - sum_{i > j} s_i v_i) = tau_j (w_j - sum_{i > j} s_i G[j][i]); one thread per vector
                if (tid < nz) {
                    double sv[RB];
#pragma unroll
                    for (int j = RB - 1; j >= 0; --j) {
                        if (j < nb) {
                            double acc = ws[tid * RB + j];
#pragma unroll
                            for (int i = j + 1; i < RB; ++i)
                                if (i < nb) acc -= sv[i] * Gs[j * RB + i];
                            sv[j] = tt[lo + j] * acc;
                            ys[tid * nb + j] = sv[j];
                        } else {
                            sv[j] = 0.0;
                        }
                    }
                }
                __syncthreads();
                // ---- C. z -= V y
                for (int e = tid; e < nz * (n - lo - 1); e += ET) {
                    const int s = e / (n - lo - 1), i = lo + 1 + (e - s * (n - lo - 1));
                    const double *y = ys + s * nb;
                    double acc = 0.0;
                    for (int j = 0; j < nb; ++j)
                        if (i >= lo + j + 1) acc += V[(size_t)j * ldv + i] * y[j];
                    zv[(size_t)s * n + i] -= acc;
                }
                __syncthreads();
            }
            cp_async_wait<0>();
            __syncthreads();
            for (int s = warp; s < nz; s += EW) {
                const int q = c + C * (s0 + s);
                double *z = zv + (size_t)s * n;
                double nn = 0.0, best = -1.0;
                int bi = 0;
                for (int i = lane; i < n; i += 32) {
                    const double v = z[i];
                    nn += v * v;
                    if (fabs(v) > best) { best = fabs(v); bi = i; }
                }
                nn = warp_sum(nn);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {               // arg-max |v|, first occurrence on ties
                    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                }
                double f = nn > 0.0 ? rsqrt(nn) : 0.0;
                if (p.flip && z[bi] < 0.0) f = -f;
                if (!(nn == nn) || !(fabs(lam[q]) <= DBL_MAX)) bad = true;
                const int colo = k - 1 - q;
                for (int i = lane; i < n; i += 32) p.U[(size_t)i * k + colo] = z[i] * f;
            }
        }
        if (p.info && bad && lane == 0) atomicExch(p.info, -1);
    }
    stamp(9);
    cluster_sync();          // peers' shared memory stays mapped until every remote access has completed
}

struct EigPlan {
    int C, ncl, kc, L, ldv, z_smem, zcap, cholqr;
    bool a_smem, f_smem;
    size_t smem_bytes, ws_bytes;
    size_t off_V, off_Z, off_lam, off_A, off_F;
};

int cluster_limit() {
    static int lim = 0;
    if (lim) return lim;
    lim = 8;
    const char *e = getenv("SKT_EIG_CLUSTER");
    int want = e ? atoi(e) : 16;
    if (want >= 16) {
        // 16-CTA clusters are a non-portable size: ask the runtime whether one fits (1 CTA per SM at this footprint)
        if (cudaFuncSetAttribute(symeig_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(16, 1, 1);
            cfg.blockDim = dim3(ET, 1, 1);
            cfg.dynamicSmemBytes = smem_optin_bytes();
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int nclus = 0;
            cudaFuncSetAttribute(symeig_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_optin_bytes());
            if (cudaOccupancyMaxActiveClusters(&nclus, symeig_cluster_kernel, &cfg) == cudaSuccess && nclus >= 1) lim = 16;
        }
        cudaGetLastError();
    } else if (want >= 1) {
        lim = want >= 8 ? 8 : (want >= 4 ? 4 : (want >= 2 ? 2 : 1));
    }
    return lim;
}

int eig_plan(int n, int k, EigPlan &pl) {
    SKT_REQUIRE(n >= 1 && n <= EIG_MAX_N, SKT_EUNSUPPORTED, "symmetric eigensolver supports 1 <= n <= %d, got %d", EIG_MAX_N, n);
    SKT_REQUIRE(k >= 1 && k <= n, SKT_EINVAL, "need 1 <= k <= n (k = %d, n = %d)", k, n);
    int C = n <= 48 ? 1 : (n <= 128 ? 4 : (n <= 256 ? 8 : 16));
    C = std::min(C, cluster_limit());
    pl.C = C;
    pl.ncl = (n + C - 1) / C;
    pl.kc = (k + C - 1) / C;
    SKT_REQUIRE(pl.kc <= ET, SKT_EUNSUPPORTED, "too many eigenvalues per CTA (%d)", pl.kc);
    int L = 1;
    while (L * 2 * pl.kc <= ET && L < 256) L *= 2;
    pl.L = L;
    pl.ldv = (n + 1) & ~1;
    const size_t limit = smem_optin_bytes();
    const EigLayout base = eig_layout(n, k, 0);
    SKT_REQUIRE(base.total_bytes + 1024 < limit, SKT_EUNSUPPORTED, "n = %d too large for the cluster eigensolver", n);
    const size_t avail = (limit - base.total_bytes) / sizeof(double) - 8;
    const size_t needA = (size_t)pl.ncl * n;
    const size_t needF = (size_t)pl.kc * fslot_doubles(n);
    const size_t needZ = (size_t)k * n;
    const size_t ringd = (size_t)2 * RB * pl.ldv;
    SKT_REQUIRE(ringd + (size_t)n <= avail, SKT_EUNSUPPORTED, "n = %d too large for the back-transformation buffers", n);
    pl.zcap = (int)std::min<size_t>(std::min<size_t>((size_t)pl.kc, ZCAP_MAX), (avail - ringd) / (size_t)n);
    const size_t need4 = ringd + (size_t)pl.zcap * n;
    pl.a_smem = needA <= avail;
    pl.f_smem = needF <= avail;
    pl.z_smem = ((pl.f_smem ? needF : 0) + needZ <= avail) ? 1 : 0;
    // the symmetric orthonormalisation of the last pass keeps its two LW x LW matrices where the LU factors were
    pl.cholqr = (pl.z_smem && pl.f_smem && needF >= (size_t)2 * LW * LW) ? 1 : 0;
    size_t bigd = need4;
    if (pl.a_smem) bigd = std::max(bigd, needA);
    bigd = std::max(bigd, (pl.f_smem ? needF : 0) + (pl.z_smem ? needZ : 0));
    pl.smem_bytes = eig_layout(n, k, bigd).total_bytes;
    size_t o = 0;
    pl.off_V = o; o += align_up((size_t)n * pl.ldv * sizeof(double), 256);
    pl.off_Z = o; o += align_up((size_t)k * n * sizeof(double), 256);
    pl.off_lam = o; o += align_up((size_t)k * sizeof(double), 256);
    pl.off_A = o; if (!pl.a_smem) o += align_up((size_t)C * needA * sizeof(double), 256);
    pl.off_F = o; if (!pl.f_smem) o += align_up((size_t)C * needF * sizeof(double), 256);
    pl.ws_bytes = o;
    return SKT_OK;
}

}  // namespace
}  // namespace skt

using namespace skt;

extern "C" size_t skt_symeig_workspace(int n, int k) {
    EigPlan pl;
    if (eig_plan(n, k, pl) != SKT_OK) return 0;
    return pl.ws_bytes;
}

extern "C" int skt_symeig_top(const double *A_d, int n, int k, int flipsign, double *W_d, double *U_d, int *info_d,
                              void *ws_d, size_t ws_bytes, skt_stream stream) {
    SKT_REQUIRE(A_d && U_d, SKT_EINVAL, "NULL argument");
    EigPlan pl;
    const int rc = eig_plan(n, k, pl);
    if (rc) return rc;
    SKT_REQUIRE(ws_d && ws_bytes >= pl.ws_bytes, SKT_EWORKSPACE, "workspace too small: need %zu bytes, got %zu", pl.ws_bytes, ws_bytes);
    SKT_REQUIRE((((uintptr_t)ws_d) & 15) == 0, SKT_EINVAL, "workspace must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    EigParams p;
    p.A = A_d; p.n = n; p.k = k; p.flip = flipsign ? 1 : 0;
    p.W = W_d; p.U = U_d; p.info = info_d;
    char *ws = (char *)ws_d;
    p.Vg = (double *)(ws + pl.off_V);
    p.Zg = (double *)(ws + pl.off_Z);
    p.lamg = (double *)(ws + pl.off_lam);
    p.Ag = pl.a_smem ? nullptr : (double *)(ws + pl.off_A);
    p.Fg = pl.f_smem ? nullptr : (double *)(ws + pl.off_F);
    p.C = pl.C; p.ncl = pl.ncl; p.kc = pl.kc; p.L = pl.L; p.ldv = pl.ldv; p.z_smem = pl.z_smem; p.zcap = pl.zcap; p.cholqr = pl.cholqr;
    if (const char *e = getenv("SKT_EIG_NO_LOEWDIN")) if (e[0] == '1') p.cholqr = 0;
    p.iters = 2;        // inverse iterations (tools/symeig_model.py: two reach the residual floor on every spectrum tried)
    if (const char *e = getenv("SKT_EIG_ITERS")) p.iters = std::max(1, atoi(e));
    p.prof = nullptr;
    const char *pe = getenv("SKT_EIG_PROF");
    const bool prof = pe && pe[0] == '1';
    if (prof) {
        SKT_CUDA(cudaMalloc((void **)&p.prof, 16 * sizeof(unsigned long long)));
        SKT_CUDA(cudaMemset(p.prof, 0, 16 * sizeof(unsigned long long)));
    }
    if (info_d) SKT_CUDA(cudaMemsetAsync(info_d, 0, sizeof(int), st));
    SKT_CUDA(cudaFuncSetAttribute(symeig_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_optin_bytes()));
    if (pl.C > 8) SKT_CUDA(cudaFuncSetAttribute(symeig_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(pl.C, 1, 1);
    cfg.blockDim = dim3(ET, 1, 1);
    cfg.dynamicSmemBytes = pl.smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = pl.C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    SKT_CUDA(cudaLaunchKernelEx(&cfg, symeig_cluster_kernel, p));
    SKT_LAUNCH_CHECK();
    if (prof) {      // debugging aid: phase timestamps of CTA 0 (synchronises)
        unsigned long long h[16];
        SKT_CUDA(cudaStreamSynchronize(st));
        SKT_CUDA(cudaMemcpy(h, p.prof, sizeof(h), cudaMemcpyDeviceToHost));
        cudaFree(p.prof);
        fprintf(stderr, "[symeig n=%d k=%d C=%d L=%d] tridiag %.1f us | eigenvalues %.1f | factor %.1f | iterations", n, k, pl.C,
                pl.L, (h[1] - h[0]) * 1e-3, (h[2] - h[1]) * 1e-3, (h[3] - h[2]) * 1e-3);
        for (int i = 0; i < p.iters && i < 5; ++i) fprintf(stderr, " %.1f", (h[4 + i] - h[3 + i]) * 1e-3);
        fprintf(stderr, " | back-transform %.1f us", (h[9] - h[3 + std::min(p.iters, 5)]) * 1e-3);
        fprintf(stderr, " || last iteration: solve %.1f, barrier %.1f, orthonormalise (CTA 0) %.1f, barrier %.1f us\n",
                (h[10] - h[2 + p.iters]) * 1e-3, (h[11] - h[10]) * 1e-3, (h[12] - h[11]) * 1e-3, (h[13] - h[12]) * 1e-3);
        fprintf(stderr, "   vectors orthonormalised symmetrically / by Gram-Schmidt on the last pass: %llu / %llu\n", h[14], h[15]);
    }
    return SKT_OK;
}

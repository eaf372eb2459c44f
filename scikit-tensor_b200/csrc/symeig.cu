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
    int C, ncl, kc, L, ldv, z_smem, iters;
    unsigned long long *prof;   // optional phase timestamps (globaltimer ns) written by CTA 0, SKT_EIG_PROF=1
};

// shared-memory layout in doubles (host and device agree through this function)
struct EigLayout {
    int vb, wb, pb, dd, ee, tt, lam, grp, red, nbs, big;
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
    L.red = o; o += 2 * EW + 2;
    L.nbs = o; o += (ET + 1) / 2 + 1;       // ints (one per multisection slot)
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

__global__ void __launch_bounds__(ET, 1) symeig_cluster_kernel(const __grid_constant__ EigParams p) {
    extern __shared__ __align__(16) double esm[];
    const int n = p.n, k = p.k, C = p.C, ncl = p.ncl, kc = p.kc;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const EigLayout Lo = eig_layout(n, k, 0);
    double *vb = esm + Lo.vb, *wb = esm + Lo.wb, *pb = esm + Lo.pb, *dd = esm + Lo.dd, *ee = esm + Lo.ee, *tt = esm + Lo.tt;
    double *lam = esm + Lo.lam, *red = esm + Lo.red, *big = esm + Lo.big;
    int *grp = reinterpret_cast<int *>(esm + Lo.grp);
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
    bool has_pend = false;
    for (int kk = 0; kk + 1 < n; ++kk) {
        const int cur = kk & 1, prev = cur ^ 1;
        double *vc = vb + cur * n, *vo = vb + prev * n, *wo = wb + prev * n;
        if (c == kk % C) {
            double *col = Ablk + (size_t)(kk / C) * n;
            const double wk = has_pend ? wo[kk] : 0.0, vk = has_pend ? vo[kk] : 0.0;
            double part = 0.0;
            for (int i = kk + tid; i < n; i += ET) {
                double a = col[i];
                if (has_pend) {
                    a -= vo[i] * wk + wo[i] * vk;
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
                tau = (beta - alpha) / beta;
                scal = 1.0 / (alpha - beta);
                ek = beta;
            }
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
        }
        cluster_sync();
        const double tau = tt[kk];
        // previous step's rank-2 update fused with this step's p = tau A v on the columns this CTA owns
        {
            const int t0 = (kk + 1 - c + C - 1) / C;      // first local column with j >= kk + 1  (j = c + C t)
            for (int t = max(t0, 0) + warp; t < ncl; t += EW) {
                const int j = c + C * t;
                if (j >= n) break;
                double *col = Ablk + (size_t)t * n;
                const double wj = has_pend ? wo[j] : 0.0, vj = has_pend ? vo[j] : 0.0;
                double acc = 0.0;
                for (int i = kk + 1 + lane; i < n; i += 32) {
                    double a = col[i];
                    if (has_pend) {
                        a -= vo[i] * wj + wo[i] * vj;
                        col[i] = a;
                    }
                    acc += a * vc[i];
                }
                acc = warp_sum(acc) * tau;
                if (lane < C) remote(pb, lane)[j] = acc;
            }
        }
        cluster_sync();
        if (tau != 0.0) {
            double part = 0.0;
            for (int i = kk + 1 + tid; i < n; i += ET) part += pb[i] * vc[i];
            const double al = -0.5 * tau * block_sum(part);
            double *wc = wb + cur * n;
            for (int i = kk + 1 + tid; i < n; i += ET) wc[i] = pb[i] + al * vc[i];
            has_pend = true;
        } else {
            has_pend = false;
        }
        __syncthreads();
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
                    const double e1 = ee[i - 1];
                    qv = dd[i] - x - e1 * e1 / qv;
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
        for (int j = 1; j < k; ++j) {
            const double sep = 10.0 * EPS * fabs(lam[j]);
            if (lam[j] - lam[j - 1] < sep) lam[j] = lam[j - 1] + sep;
            grp[j] = (lam[j] - lam[j - 1] < ortol) ? grp[j - 1] : j;
        }
    }
    __syncthreads();

    // ================================================================== 3. eigenvectors of T (inverse iteration)
    const int fslot = fslot_doubles(n);
    double *Freg = p.Fg ? p.Fg + (size_t)c * kc * fslot : big;
    double *Zs = p.z_smem ? big + (p.Fg ? 0 : (size_t)kc * fslot) : p.Zg;     // CTA 0's orthogonalisation copy
    const int myk = (k > c) ? (k - c + C - 1) / C : 0;                        // eigenvalues owned by this CTA: q = c + C s
    double alast = 0.0;
    if (tid < myk) {
        // ---- dlagtf: T - lambda I = P L U, one thread per eigenvalue
        const int q = c + C * tid;
        double *ra = Freg + (size_t)tid * fslot, *bb = ra + n, *cc = bb + n, *d2 = cc + n;
        unsigned char *sw = reinterpret_cast<unsigned char *>(d2 + 2 * n);
        const double lq = lam[q];
        double ai = dd[0] - lq, bi = n > 1 ? ee[0] : 0.0, amax = 0.0;
        double scale1 = fabs(ai) + fabs(bi);
        for (int i = 0; i + 1 < n; ++i) {
            const double ci = ee[i], a1 = dd[i + 1] - lq, b1 = (i + 2 < n) ? ee[i + 1] : 0.0;
            const double scale2 = fabs(ci) + fabs(a1) + fabs(b1);
            const bool noswap = (ci == 0.0) || (ai != 0.0 && fabs(ci) * scale1 <= fabs(ai) * scale2);
            double m, ui, bo, di, an, bn;
            if (noswap) {
                m = (ci == 0.0) ? 0.0 : ci / ai;
                ui = ai; bo = bi; di = 0.0;
                an = a1 - m * bi; bn = b1;
            } else {
                m = ai / ci;
                ui = ci; bo = a1; di = b1;
                an = bi - m * a1; bn = -m * b1;
            }
            ra[i] = ui; bb[i] = bo; cc[i] = m; d2[i] = di; sw[i] = noswap ? 0 : 1;
            amax = fmax(amax, fmax(fabs(ui), fmax(fabs(bo), fabs(di))));
            ai = an; bi = bn; scale1 = scale2;
        }
        ra[n - 1] = ai;
        amax = fmax(amax, fabs(ai));
        alast = fabs(ai);
        const double tol = (amax == 0.0) ? EPS : fmax(amax * EPS, SAFMIN / EPS);
        // pivots below tol are replaced by +-tol (dlagts, job = -1); the reciprocals are taken by a whole warp below
        d2[n - 1] = tol;                        // (d2 has n slots, U uses the first n - 2): parked for that pass
    }
    __syncthreads();
    // reciprocal pivots and start vectors: every thread, warp per eigenvalue
    for (int s = warp; s < myk; s += EW) {
        const int q = c + C * s;
        double *ra = Freg + (size_t)s * fslot, *d2 = ra + 3 * n, *z = ra + 4 * n;
        const double tol = d2[n - 1];
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
            const double *z = Freg + (size_t)s * fslot + 4 * n;
            double a = 0.0;
            for (int i = lane; i < n; i += 32) a += fabs(z[i]);
            a = warp_sum(a);
            if (lane == 0) wb[s] = a;            // wb (2n doubles, free after phase 1) as per-eigenvalue scratch: kc <= n
        }
        __syncthreads();
        if (tid < myk) {
            double *ra = Freg + (size_t)tid * fslot, *bb = ra + n, *cc = bb + n, *d2 = cc + n, *z = d2 + n;
            const unsigned char *sw = reinterpret_cast<const unsigned char *>(d2 + 2 * n);
            const double nrm1 = wb[tid];
            const double scl = n * onenrm * fmax(EPS, alast) / (nrm1 == 0.0 ? 1.0 : nrm1);
            // forward: apply P L
            double yp = z[0] * scl;
            for (int i = 1; i < n; ++i) {
                const double yi = z[i] * scl, m = cc[i - 1];
                if (sw[i - 1]) {
                    z[i - 1] = yi;
                    yp = yp - m * yi;
                } else {
                    z[i - 1] = yp;
                    yp = yi - m * yp;
                }
            }
            z[n - 1] = yp;
            // backward: U x = y
            double y1 = 0.0, y2 = 0.0, zmax = 0.0;
            for (int i = n - 1; i >= 0; --i) {
                double t = z[i];
                if (i + 1 < n) t -= bb[i] * y1;
                if (i + 2 < n) t -= d2[i] * y2;
                t *= ra[i];
                z[i] = t;
                y2 = y1; y1 = t;
                zmax = fmax(zmax, fabs(t));
            }
            wb[tid] = (zmax > 0.0 && zmax < DBL_MAX) ? 1.0 / zmax : 1.0;
        }
        __syncthreads();
        for (int s = warp; s < myk; s += EW) {
            const int q = c + C * s;
            const double *z = Freg + (size_t)s * fslot + 4 * n;
            const double sc = wb[s];
            double *zg = p.Zg + (size_t)q * n;
            for (int i = lane; i < n; i += 32) zg[i] = z[i] * sc;
        }
        __threadfence();
        cluster_sync();
        if (c == 0) {
            // ---- right-looking modified Gram-Schmidt inside the groups, ascending; then unit 2-norm
            // (without the shared copy the vectors are read where the peers wrote them: L2 loads, never a stale L1 line)
            auto ldz = [&](const double *q_) -> double { return p.z_smem ? *q_ : __ldcg(q_); };
            if (p.z_smem) {
                for (size_t e = tid; e < (size_t)k * n; e += ET) Zs[e] = __ldcg(p.Zg + e);
                __syncthreads();
            }
            for (int i = 0; i + 1 < k; ++i) {
                if (grp[i + 1] != grp[i]) continue;            // z_i is the last of its group: nothing to project
                const double *zi = Zs + (size_t)i * n;
                double nn = 0.0;
                for (int r = lane; r < n; r += 32) { const double v = ldz(zi + r); nn += v * v; }
                nn = warp_sum(nn);
                for (int j = i + 1 + warp; j < k && grp[j] == grp[i]; j += EW) {
                    double *zj = Zs + (size_t)j * n;
                    double dot = 0.0;
                    for (int r = lane; r < n; r += 32) dot += ldz(zi + r) * ldz(zj + r);
                    dot = warp_sum(dot);
                    const double f = nn > 0.0 ? dot / nn : 0.0;
                    for (int r = lane; r < n; r += 32) zj[r] = ldz(zj + r) - f * ldz(zi + r);
                }
                __syncthreads();
            }
            for (int j = warp; j < k; j += EW) {
                double *zj = Zs + (size_t)j * n;
                double nn = 0.0;
                for (int r = lane; r < n; r += 32) { const double v = ldz(zj + r); nn += v * v; }
                nn = warp_sum(nn);
                const double f = nn > 0.0 ? rsqrt(nn) : 0.0;
                for (int r = lane; r < n; r += 32) {
                    const double v = ldz(zj + r) * f;
                    zj[r] = v;
                    if (p.z_smem) p.Zg[(size_t)j * n + r] = v;
                }
            }
        }
        __threadfence();
        cluster_sync();
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
    {
        // big region: [kc][n] vectors, then the reflector ring [2][RB][ldv]
        const int ldv = p.ldv;
        double *zv = big;
        double *ring = big + (((size_t)kc * n + 1) & ~(size_t)1);
        __syncthreads();
        for (int s = 0; s < kc; ++s) {
            const int q = c + C * s;
            if (q >= k) break;
            for (int i = tid; i < n; i += ET) zv[(size_t)s * n + i] = __ldcg(p.Zg + (size_t)q * n + i);
        }
        __syncthreads();
        const int nref = n - 2;                          // reflectors 0 .. n-3 (the last step's is the identity)
        const int nchunk = nref > 0 ? (nref + RB - 1) / RB : 0;
        auto issue_chunk = [&](int ch) {                 // chunk ch holds reflectors hi .. hi-RB+1, hi = nref-1 - ch*RB
            double *dst = ring + (size_t)(ch & 1) * RB * ldv;
            const int hi = nref - 1 - ch * RB;
            const int per = ldv / 2;                     // 16-byte pieces per column
            for (int e = tid; e < RB * per; e += ET) {
                const int b = e / per, piece = e - b * per;
                const int kk = hi - b;
                if (kk >= 0) cp_async16(dst + (size_t)b * ldv + 2 * piece, p.Vg + (size_t)kk * ldv + 2 * piece, 16);
            }
            cp_async_commit();
        };
        if (nchunk > 0) issue_chunk(0);
        for (int ch = 0; ch < nchunk; ++ch) {
            if (ch + 1 < nchunk) issue_chunk(ch + 1); else cp_async_commit();
            cp_async_wait<1>();
            __syncthreads();
            const double *src = ring + (size_t)(ch & 1) * RB * ldv;
            const int hi = nref - 1 - ch * RB;
            for (int s = warp; s < kc; s += EW) {
                if (c + C * s >= k) break;
                double *z = zv + (size_t)s * n;
                for (int b = 0; b < RB; ++b) {
                    const int kk = hi - b;
                    if (kk < 0) break;
                    const double tau = tt[kk];
                    if (tau == 0.0) continue;
                    const double *v = src + (size_t)b * ldv;
                    double dot = 0.0;
                    for (int i = kk + 1 + lane; i < n; i += 32) dot += v[i] * z[i];
                    dot = warp_sum(dot) * tau;
                    for (int i = kk + 1 + lane; i < n; i += 32) z[i] -= dot * v[i];
                    __syncwarp();
                }
            }
            __syncthreads();
        }
        cp_async_wait<0>();
        bool bad = false;
        for (int s = warp; s < kc; s += EW) {
            const int q = c + C * s;
            if (q >= k) break;
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
        if (p.info && bad && lane == 0) atomicExch(p.info, -1);
    }
    stamp(9);
    cluster_sync();          // peers' shared memory stays mapped until every remote access has completed
}

struct EigPlan {
    int C, ncl, kc, L, ldv, z_smem;
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
    const size_t need4 = (((size_t)pl.kc * n + 1) & ~(size_t)1) + (size_t)2 * RB * pl.ldv;
    SKT_REQUIRE(need4 <= avail, SKT_EUNSUPPORTED, "n = %d, k = %d too large for the back-transformation buffers", n, k);
    pl.a_smem = needA <= avail;
    pl.f_smem = needF <= avail;
    pl.z_smem = ((pl.f_smem ? needF : 0) + needZ <= avail) ? 1 : 0;
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
    p.C = pl.C; p.ncl = pl.ncl; p.kc = pl.kc; p.L = pl.L; p.ldv = pl.ldv; p.z_smem = pl.z_smem;
    p.iters = 3;
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
        fprintf(stderr, " | back-transform %.1f us\n", (h[9] - h[3 + std::min(p.iters, 5)]) * 1e-3);
    }
    return SKT_OK;
}

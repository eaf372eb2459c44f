// Micro-benchmark (development aid): how many warps / how much LDS feeding the FP64 tensor pipe tolerates.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int LDSFED>
__global__ void probe(double *out, int iters) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 1.0 + i * 1e-9;
    __syncthreads();
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double a0 = 1.0 + lane * 1e-9, a1 = a0, b = 1.0 - lane * 1e-9;
    for (int it = 0; it < iters; ++it) {
        if (LDSFED) {
            // 2 A loads + 4 B loads per 8 DMMAs, like one s4 step of the MTTKRP kernel
            const double *p = sm + ((it & 15) * 256 + warp * 8 + lane) % 3800;
            a0 = p[0]; a1 = p[32];
            double bb[4];
#pragma unroll
            for (int nn = 0; nn < 4; ++nn) bb[nn] = p[64 + 32 * nn];
#pragma unroll
            for (int nn = 0; nn < 4; ++nn) { dmma884(c[2 * nn][0], c[2 * nn][1], a0, bb[nn]); dmma884(c[2 * nn + 1][0], c[2 * nn + 1][1], a1, bb[nn]); }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a0, b);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}
int main() {
    double *d; cudaMalloc(&d, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 20000;
    for (int fed = 0; fed < 2; ++fed)
        for (int warps : {4, 8, 12, 16, 20, 24, 32}) {
            float best = 1e30f;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                if (fed) probe<1><<<sms, warps * 32, 32768>>>(d, iters); else probe<0><<<sms, warps * 32, 32768>>>(d, iters);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep) best = ms < best ? ms : best;
            }
            const double flops = 2.0 * 256.0 * 8.0 * iters * warps * (double)sms;
            printf("%s  %2d warps/SM: %.1f TFLOP/s\n", fed ? "LDS-fed (6 LDS per 8 DMMA)" : "register-only", warps, flops / (best * 1e-3) / 1e12);
        }
    return 0;
}

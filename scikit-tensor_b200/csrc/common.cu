// Error state, device queries and the FP64 peak probes of libsktensor_b200.
#include <stdarg.h>
#include <algorithm>
#include <string.h>

#include "common.cuh"

namespace skt {

static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    const char *base = strrchr(file, '/');
    set_error("CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, base ? base + 1 : file, line);
    return SKT_ECUDA;
}

static int g_reserved_sms = 0;

int sm_count_total();

int sm_count() { return std::max(1, sm_count_total() - g_reserved_sms); }

// device properties are cached per device (a process may switch devices between calls)
static int current_device_slot() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return -1; }
    return (dev >= 0 && dev < 64) ? dev : -1;
}

bool first_use_on_device(unsigned long long &mask) {
    const int slot = current_device_slot();
    if (slot < 0) return true;
    const unsigned long long bit = 1ULL << slot;
    if (mask & bit) return false;
    mask |= bit;
    return true;
}

int sm_count_total() {
    static int cache[64] = {0};
    const int slot = current_device_slot();
    if (slot >= 0 && cache[slot] > 0) return cache[slot];
    int n = 0;
    if (slot < 0 || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, slot) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return 148;  // B200; only reached when no device is visible (workspace queries on a CPU box)
    }
    cache[slot] = n;
    return n;
}

size_t smem_optin_bytes() {
    static int cache[64] = {0};
    const int slot = current_device_slot();
    if (slot >= 0 && cache[slot] > 0) return (size_t)cache[slot];
    int v = 0;
    if (slot < 0 || cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, slot) != cudaSuccess || v <= 0) {
        cudaGetLastError();
        return 232448;
    }
    cache[slot] = v;
    return (size_t)v;
}

namespace {

// 8 independent DMMA chains per warp; measures the FP64 tensor-pipe issue rate.
__global__ void probe_dmma_kernel(double *out, int iters) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

// 16 independent DFMA chains per thread; measures the FP64 FMA issue rate.
__global__ void probe_dfma_kernel(double *out, int iters) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = i * 1e-3;
    const double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 123.456) out[0] = s;
}

}  // namespace
}  // namespace skt

using namespace skt;

extern "C" int skt_version(void) { return 100; }

extern "C" uint64_t skt_launch_count(void) { return g_launches; }

extern "C" const char *skt_last_error(void) { return g_err; }

extern "C" int skt_device_info(int *sm, int *cc_major, int *cc_minor, size_t *smem_optin) {
    int dev = 0;
    SKT_CUDA(cudaGetDevice(&dev));
    int v = 0;
    if (sm) { SKT_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev)); *sm = v; }
    if (cc_major) { SKT_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev)); *cc_major = v; }
    if (cc_minor) { SKT_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev)); *cc_minor = v; }
    if (smem_optin) { SKT_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)); *smem_optin = (size_t)v; }
    return SKT_OK;
}

extern "C" int skt_probe_fp64_peaks(double *dmma_tflops, double *dfma_tflops) {
    double *d = nullptr;
    SKT_CUDA(cudaMalloc(&d, 8));
    cudaEvent_t e0, e1;
    SKT_CUDA(cudaEventCreate(&e0));
    SKT_CUDA(cudaEventCreate(&e1));
    const int sms = sm_count_total(), blocks = sms * 4, threads = 256, iters = 20000;
    float ms = 0.f;
    double best;
    if (dmma_tflops) {
        best = 0.0;
        for (int rep = 0; rep < 4; ++rep) {
            SKT_CUDA(cudaEventRecord(e0));
            probe_dmma_kernel<<<blocks, threads>>>(d, iters);
            SKT_CUDA(cudaEventRecord(e1));
            SKT_CUDA(cudaEventSynchronize(e1));
            SKT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            const double flops = 2.0 * 256.0 * 8.0 * iters * (threads / 32) * (double)blocks;
            if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
        }
        *dmma_tflops = best;
    }
    if (dfma_tflops) {
        best = 0.0;
        for (int rep = 0; rep < 4; ++rep) {
            SKT_CUDA(cudaEventRecord(e0));
            probe_dfma_kernel<<<blocks, threads>>>(d, iters);
            SKT_CUDA(cudaEventRecord(e1));
            SKT_CUDA(cudaEventSynchronize(e1));
            SKT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            const double flops = 2.0 * 16.0 * iters * threads * (double)blocks;
            if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
        }
        *dfma_tflops = best;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return SKT_OK;
}

extern "C" int skt_set_reserved_sms(int n) {
    const int prev = skt::g_reserved_sms;
    skt::g_reserved_sms = (n < 0) ? 0 : n;
    return prev;
}

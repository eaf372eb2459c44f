// Latency probe for the primitives of the cluster eigensolver's inner loop (B200, sm_100a):
// cluster barrier (release/acquire) alone, with DSMEM broadcast stores in front, block reductions, warp reductions.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/cluster_probe tools/cluster_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void csync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void csync_relaxed() {
    asm volatile("barrier.cluster.arrive.relaxed.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.aligned;\n" ::: "memory");
}
__device__ __forceinline__ double *remote(double *ptr, unsigned rank) {
    unsigned long long in = (unsigned long long)ptr, out;
    asm volatile("mapa.u64 %0, %1, %2;\n" : "=l"(out) : "l"(in), "r"(rank));
    return (double *)out;
}
__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void probe(int mode, int iters, int C, unsigned long long *out, double *sink) {
    __shared__ double buf[2048];
    __shared__ double red[32];
    unsigned crank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(crank));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    buf[tid] = tid;
    csync();
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t0));
    double acc = 0.0;
    for (int it = 0; it < iters; ++it) {
        if (mode == 0) {
            csync();
        } else if (mode == 1) {
            csync_relaxed();
        } else if (mode == 2) {            // owner broadcasts 1 value per thread to every CTA, then barrier
            if ((int)crank == it % C)
                for (int r = 0; r < C; ++r) remote(buf, r)[tid] = acc + it;
            csync();
            acc += buf[(tid + it) & 511];
        } else if (mode == 3) {            // block reduction (2 __syncthreads) only
            double v = warp_sum(buf[tid] + acc);
            __syncthreads();
            if (lane == 0) red[warp] = v;
            __syncthreads();
            double s = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
            acc += s * 1e-30;
        } else if (mode == 4) {            // warp reduction chain
            acc += warp_sum(buf[tid] + acc) * 1e-30;
        } else if (mode == 5) {            // every CTA pushes 1 value per lane<C of each warp, then barrier
            if (lane < C) remote(buf, lane)[warp + 32 * (int)crank] = acc;
            csync();
            acc += buf[(tid + it) & 511];
        } else if (mode == 6) {            // pull: barrier, then every thread loads one value from the owner's shared memory
            if ((int)crank == it % C) buf[tid] = acc + it;
            csync();
            acc += remote(buf, it % C)[tid];
        }
    }
    asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t1));
    if (tid == 0 && crank == 0) out[0] = t1 - t0;
    sink[blockIdx.x * blockDim.x + tid] = acc;
    csync();
}

int main() {
    unsigned long long *out;
    double *sink;
    cudaMalloc(&out, 8);
    cudaMalloc(&sink, 16 * 512 * 8);
    cudaFuncSetAttribute(probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    const char *names[] = {"cluster barrier release/acquire", "cluster barrier relaxed", "owner broadcast (C stores/thread) + barrier",
                           "block reduction (2 __syncthreads)", "warp reduction (5 shuffles, fp64)",
                           "all-gather push (lanes < C) + barrier", "owner local store + barrier + DSMEM pull"};
    for (int C : {1, 2, 4, 8, 16}) {
        for (int mode = 0; mode < 7; ++mode) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(C);
            cfg.blockDim = dim3(512);
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            const int iters = 2000;
            for (int rep = 0; rep < 2; ++rep) {
                cudaError_t e = cudaLaunchKernelEx(&cfg, probe, mode, iters, C, out, sink);
                if (e != cudaSuccess) { printf("C=%d launch failed: %s\n", C, cudaGetErrorString(e)); break; }
                cudaDeviceSynchronize();
            }
            unsigned long long ns = 0;
            cudaMemcpy(&ns, out, 8, cudaMemcpyDeviceToHost);
            printf("C=%2d  %-48s %8.1f ns per iteration\n", C, names[mode], (double)ns / iters);
        }
    }
    return 0;
}

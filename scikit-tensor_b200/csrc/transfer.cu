// Host -> device ingest of a user's (pageable) tensor buffer.
//
// The reference keeps X in host memory and every call starts from there (dtensor is an ndarray view,
// sktensor/dtensor.py:48-50), so the first thing cp_als / tucker_hooi / uttkrp do here is move X to HBM.
// A plain cudaMemcpy from pageable memory stages through one driver-owned bounce buffer on one thread
// (~10 GB/s); this path stages through a pool of pinned buffers with several host threads, each
// pipelining memcpy (pageable -> pinned) against its own asynchronous DMA (pinned -> HBM), which is
// bounded by the PCIe link instead.  The copy is ordered on the caller's stream: work enqueued on
// `stream` after skt_upload sees the data; the host buffer is no longer referenced once skt_upload returns.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

#if defined(__x86_64__) || defined(__SSE2__)
#include <emmintrin.h>
#define SKT_HAVE_SSE2 1
#endif

namespace skt {
namespace {

// pageable -> pinned staging copy with non-temporal stores: the staging buffer is read next by the DMA engine, not by
// the CPU, so writing it through the cache (read-for-ownership + eviction) only costs memory bandwidth
inline void stage_copy(void *dst, const void *src, size_t n) {
#ifdef SKT_HAVE_SSE2
    if ((((uintptr_t)dst) & 15) == 0 && n >= 256) {
        const char *s = (const char *)src;
        char *d = (char *)dst;
        size_t i = 0;
        for (; i + 64 <= n; i += 64) {
            const __m128i a = _mm_loadu_si128((const __m128i *)(s + i));
            const __m128i b = _mm_loadu_si128((const __m128i *)(s + i + 16));
            const __m128i c = _mm_loadu_si128((const __m128i *)(s + i + 32));
            const __m128i e = _mm_loadu_si128((const __m128i *)(s + i + 48));
            _mm_stream_si128((__m128i *)(d + i), a);
            _mm_stream_si128((__m128i *)(d + i + 16), b);
            _mm_stream_si128((__m128i *)(d + i + 32), c);
            _mm_stream_si128((__m128i *)(d + i + 48), e);
        }
        _mm_sfence();
        if (i < n) memcpy(d + i, s + i, n - i);
        return;
    }
#endif
    memcpy(dst, src, n);
}

constexpr size_t CHUNK = 4u << 20;   // bytes per staging buffer
constexpr int MAX_THREADS = 16;
constexpr int SLOTS = 2;             // staging buffers per thread (memcpy of one overlaps the DMA of the other)

struct Pool {
    int device = -1;
    int nthreads = 0;
    char *pin[MAX_THREADS][SLOTS] = {};
    cudaEvent_t ev[MAX_THREADS][SLOTS] = {};
    cudaEvent_t done[MAX_THREADS] = {};
    cudaEvent_t start = nullptr;
    cudaStream_t cs[MAX_THREADS] = {};
};

std::mutex g_mu;
Pool g_pool;

int pool_init(Pool &P, int device) {
    if (P.device == device && P.nthreads > 0) return SKT_OK;
    if (P.nthreads > 0) {   // device changed: rebuild
        for (int t = 0; t < P.nthreads; ++t) {
            for (int s = 0; s < SLOTS; ++s) { cudaFreeHost(P.pin[t][s]); cudaEventDestroy(P.ev[t][s]); }
            cudaEventDestroy(P.done[t]);
            cudaStreamDestroy(P.cs[t]);
        }
        cudaEventDestroy(P.start);
        P = Pool();
    }
    // a share of the host cores, divided fairly between the ranks of one node (torchrun exports LOCAL_WORLD_SIZE)
    unsigned hc = std::thread::hardware_concurrency();
    unsigned ranks = 1;
    if (const char *lw = getenv("LOCAL_WORLD_SIZE")) ranks = (unsigned)std::max(1, atoi(lw));
    // 5/8 of the cores: measured plateau at the PCIe limit (16-core host: 8 threads 45 GB/s, 10..16 threads 49-50 GB/s)
    int nt = (int)std::max(1u, std::min<unsigned>(MAX_THREADS, (hc ? hc * 5 / 8 : 4) / ranks));
    if (ranks > 1) nt = std::max(nt, 2);
    const char *env = getenv("SKT_UPLOAD_THREADS");
    if (env && atoi(env) > 0) nt = std::min(MAX_THREADS, atoi(env));
    for (int t = 0; t < nt; ++t) {
        for (int s = 0; s < SLOTS; ++s) {
            SKT_CUDA(cudaHostAlloc((void **)&P.pin[t][s], CHUNK, cudaHostAllocDefault));
            SKT_CUDA(cudaEventCreateWithFlags(&P.ev[t][s], cudaEventDisableTiming));
        }
        SKT_CUDA(cudaEventCreateWithFlags(&P.done[t], cudaEventDisableTiming));
        SKT_CUDA(cudaStreamCreateWithFlags(&P.cs[t], cudaStreamNonBlocking));
    }
    SKT_CUDA(cudaEventCreateWithFlags(&P.start, cudaEventDisableTiming));
    P.device = device;
    P.nthreads = nt;
    return SKT_OK;
}

}  // namespace
}  // namespace skt

using namespace skt;

extern "C" int skt_upload(void *dst_d, const void *src_h, size_t nbytes, skt_stream stream) {
    SKT_REQUIRE((dst_d && src_h) || nbytes == 0, SKT_EINVAL, "NULL argument");
    if (nbytes == 0) return SKT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (nbytes < 2 * CHUNK) {   // small: one plain copy
        SKT_CUDA(cudaMemcpyAsync(dst_d, src_h, nbytes, cudaMemcpyHostToDevice, st));
        SKT_CUDA(cudaStreamSynchronize(st));
        return SKT_OK;
    }
    {
        // page-locked source (cudaHostAlloc / cudaHostRegister / torch pin_memory): the DMA engine reads it in place --
        // no staging copy, no host threads; N ranks of one node then share only the PCIe links, not the host cores
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, src_h) == cudaSuccess && attr.type == cudaMemoryTypeHost) {
            SKT_CUDA(cudaMemcpyAsync(dst_d, src_h, nbytes, cudaMemcpyHostToDevice, st));
            SKT_CUDA(cudaStreamSynchronize(st));
            return SKT_OK;
        }
        cudaGetLastError();      // (pageable memory reports an error on older drivers: not ours)
    }
    std::lock_guard<std::mutex> lock(g_mu);
    int device = 0;
    SKT_CUDA(cudaGetDevice(&device));
    int rc = pool_init(g_pool, device);
    if (rc != SKT_OK) return rc;
    Pool &P = g_pool;
    const size_t nchunks = (nbytes + CHUNK - 1) / CHUNK;
    const int nt = (int)std::min<size_t>(P.nthreads, nchunks);
    // the copy streams start after whatever the caller's stream has queued so far (dst may be recycled memory)
    SKT_CUDA(cudaEventRecord(P.start, st));
    std::vector<cudaError_t> err(nt, cudaSuccess);
    auto worker = [&](int t) {
        cudaError_t e = cudaSetDevice(device);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(P.cs[t], P.start, 0);
        int it = 0;
        for (size_t c = t; c < nchunks && e == cudaSuccess; c += nt, ++it) {
            const int s = it % SLOTS;
            if (it >= SLOTS) e = cudaEventSynchronize(P.ev[t][s]);   // the slot's previous DMA has drained
            if (e != cudaSuccess) break;
            const size_t off = c * CHUNK, len = std::min(CHUNK, nbytes - off);
            stage_copy(P.pin[t][s], (const char *)src_h + off, len);
            e = cudaMemcpyAsync((char *)dst_d + off, P.pin[t][s], len, cudaMemcpyHostToDevice, P.cs[t]);
            if (e == cudaSuccess) e = cudaEventRecord(P.ev[t][s], P.cs[t]);
        }
        if (e == cudaSuccess) e = cudaEventRecord(P.done[t], P.cs[t]);
        err[t] = e;
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(worker, t);
    worker(0);
    for (auto &x : th) x.join();
    for (int t = 0; t < nt; ++t)
        if (err[t] != cudaSuccess) return cuda_fail(err[t], "skt_upload worker", __FILE__, __LINE__);
    for (int t = 0; t < nt; ++t) SKT_CUDA(cudaStreamWaitEvent(st, P.done[t], 0));
    // the staging buffers are reused by the next call: drain the DMAs before handing control back
    for (int t = 0; t < nt; ++t) SKT_CUDA(cudaEventSynchronize(P.done[t]));
    return SKT_OK;
}

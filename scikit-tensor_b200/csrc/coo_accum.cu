// De-duplication of COO coordinates on the device -- the `accumfun` path of sptensor.__init__
// (reference: sktensor/sptensor.py:80-84 -> sktensor/utils.py:5-26 `accum`: lexsort of the subscripts (last mode most
// significant), boundaries where any subscript changes, `func` over each run of values).  Here: one 64-bit key per
// non-zero (the column-major linear index), one stable radix sort (own kernels, radix_sort.cuh), run heads by comparison
// with the predecessor, an inclusive scan for the output slots, and one thread per run reducing its values in storage
// order (sum, max or min -- the reducers that can run on the device; other callables stay on the host).
// The same machinery numbers the distinct keys of an array (skt_compact_keys): the sparse unfolding X_(n) only has as
// many columns as there are distinct (i_m)_{m != n} combinations, which is what the sparse nvecs operator needs.
#include <algorithm>
#include <cstring>

#include "common.cuh"
#include "radix_sort.cuh"

namespace skt {
namespace {

__global__ void accum_key_kernel(const long long *const *__restrict__ subs, const long long *__restrict__ base,
                                 const long long *__restrict__ stride, int ndim, long long n, unsigned long long *key,
                                 unsigned int *perm) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        unsigned long long k = 0;
        for (int m = 0; m < ndim; ++m) k += (unsigned long long)(subs[m][i] - base[m]) * (unsigned long long)stride[m];
        key[i] = k;
        perm[i] = (unsigned int)i;
    }
}
// slot[i] = 1 where a run of equal keys starts (exclusive-scanned in place afterwards: slot[i] = runs before i)
__global__ void head_flag_kernel(const unsigned long long *__restrict__ key, long long n, unsigned int *slot) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        slot[i] = (i == 0 || key[i] != key[i - 1]) ? 1u : 0u;
}
// one thread per run head: reduce the run's values in storage order
__global__ void run_reduce_kernel(const unsigned long long *__restrict__ key, const unsigned int *__restrict__ perm,
                                  const unsigned int *__restrict__ slot, const double *__restrict__ vals, long long n, int op,
                                  unsigned long long *out_key, double *out_val) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (i > 0 && key[i] == key[i - 1]) continue;
        const unsigned long long k = key[i];
        double acc = vals[perm[i]];
        for (long long j = i + 1; j < n && key[j] == k; ++j) {
            const double v = vals[perm[j]];
            acc = (op == 0) ? acc + v : (op == 1 ? (v > acc || v != v ? v : acc) : (v < acc || v != v ? v : acc));
        }
        out_key[slot[i]] = k;          // exclusive count of run heads before i = this run's output slot
        out_val[slot[i]] = acc;
    }
}
// id of every element's key among the sorted distinct keys, scattered back to the original order
__global__ void scatter_ids_kernel(const unsigned long long *__restrict__ key, const unsigned int *__restrict__ perm,
                                   const unsigned int *__restrict__ slot, long long n, int *ids) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const bool head = (i == 0 || key[i] != key[i - 1]);
        ids[perm[i]] = (int)slot[i] - (head ? 0 : 1);       // runs started up to and including i, minus one
    }
}

__global__ void iota_u32_kernel(unsigned int *p, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        p[i] = (unsigned int)i;
}

int blocks_for(long long n) { return (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, 8LL * sm_count())); }

struct SortedKeys {
    unsigned long long *key = nullptr, *key_b = nullptr, *key_sorted = nullptr;
    unsigned int *perm = nullptr, *perm_b = nullptr, *perm_sorted = nullptr;
    unsigned int *slot = nullptr;
    void *tmp = nullptr;
    void release() { cudaFree(key); cudaFree(key_b); cudaFree(perm); cudaFree(perm_b); cudaFree(slot); cudaFree(tmp); }
};

// sort `key` (already filled, with perm = iota), flag run heads, exclusive scan -> slot; returns the number of distinct keys
int sort_and_number(SortedKeys &S, long long n, int bits, cudaStream_t st, long long *ndistinct) {
    SKT_CUDA(cudaMalloc(&S.tmp, rsort::scratch_bytes(n)));
    int where = 0;
    const int rc = rsort::sort_pairs<unsigned long long>(S.key, S.perm, S.key_b, S.perm_b, n, bits, S.tmp, st, &where);
    if (rc != SKT_OK) return rc;
    S.key_sorted = where ? S.key_b : S.key;
    S.perm_sorted = where ? S.perm_b : S.perm;
    head_flag_kernel<<<blocks_for(n), 256, 0, st>>>(S.key_sorted, n, S.slot);
    SKT_LAUNCH_CHECK();
    unsigned int last_flag = 0, last_excl = 0;
    SKT_CUDA(cudaMemcpyAsync(&last_flag, S.slot + (n - 1), sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    const int rc2 = rsort::inclusive_scan_u32(S.slot, n, S.tmp, st);       // (exclusive, in place)
    if (rc2 != SKT_OK) return rc2;
    SKT_CUDA(cudaMemcpyAsync(&last_excl, S.slot + (n - 1), sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    SKT_CUDA(cudaStreamSynchronize(st));
    *ndistinct = (long long)last_excl + last_flag;
    return SKT_OK;
}

int alloc_sorted(SortedKeys &S, long long n) {
    SKT_CUDA(cudaMalloc(&S.key, n * sizeof(unsigned long long)));
    SKT_CUDA(cudaMalloc(&S.key_b, n * sizeof(unsigned long long)));
    SKT_CUDA(cudaMalloc(&S.perm, n * sizeof(unsigned int)));
    SKT_CUDA(cudaMalloc(&S.perm_b, n * sizeof(unsigned int)));
    SKT_CUDA(cudaMalloc(&S.slot, n * sizeof(unsigned int)));
    return SKT_OK;
}

}  // namespace
}  // namespace skt

using namespace skt;

extern "C" int skt_coo_accumulate_host(const int64_t *const *subs_h, const double *vals_h, int64_t nnz, const int64_t *base,
                                       const int64_t *extent, int ndim, int op, uint64_t *out_keys_h, double *out_vals_h,
                                       int64_t *out_nnz) {
    SKT_REQUIRE(subs_h && vals_h && base && extent && out_keys_h && out_vals_h && out_nnz, SKT_EINVAL, "NULL argument");
    SKT_REQUIRE(ndim >= 1 && ndim <= SKT_MAX_NDIM, SKT_EINVAL, "ndim must be in [1, %d]", SKT_MAX_NDIM);
    SKT_REQUIRE(nnz >= 1 && nnz < (1LL << 31) - 1, SKT_ESHAPE, "nnz %lld outside [1, 2^31)", (long long)nnz);
    SKT_REQUIRE(op >= 0 && op <= 2, SKT_EINVAL, "op must be 0 (sum), 1 (max) or 2 (min)");
    long long stride[SKT_MAX_NDIM];
    long double total = 1.0L;
    int bits = 1;
    {
        unsigned long long acc = 1;
        for (int m = 0; m < ndim; ++m) {
            SKT_REQUIRE(extent[m] >= 1, SKT_ESHAPE, "extent[%d] must be positive", m);
            stride[m] = (long long)acc;
            total *= (long double)extent[m];
            SKT_REQUIRE(total < 9.0e18L, SKT_ESHAPE, "the linear index of the coordinates does not fit 63 bits");
            acc *= (unsigned long long)extent[m];
        }
        while (bits < 64 && (1ULL << bits) < acc) ++bits;
    }
    cudaStream_t st = nullptr;
    long long *subs_d[SKT_MAX_NDIM] = {nullptr};
    long long **subs_dd = nullptr, *base_d = nullptr, *stride_d = nullptr;
    double *vals_d = nullptr, *oval_d = nullptr;
    unsigned long long *okey_d = nullptr;
    SortedKeys S;
    int rc = SKT_OK;
    auto cleanup = [&]() {
        for (int m = 0; m < ndim; ++m) cudaFree(subs_d[m]);
        cudaFree(subs_dd); cudaFree(base_d); cudaFree(stride_d); cudaFree(vals_d); cudaFree(oval_d); cudaFree(okey_d);
        S.release();
    };
#define ACC_TRY(expr)                          \
    do {                                       \
        rc = (expr);                           \
        if (rc != SKT_OK) { cleanup(); return rc; } \
    } while (0)
#define ACC_CUDA(call)                                                         \
    do {                                                                       \
        cudaError_t e__ = (call);                                              \
        if (e__ != cudaSuccess) { rc = cuda_fail(e__, #call, __FILE__, __LINE__); cleanup(); return rc; } \
    } while (0)
    for (int m = 0; m < ndim; ++m) {
        ACC_CUDA(cudaMalloc(&subs_d[m], nnz * sizeof(long long)));
        ACC_TRY(skt_upload(subs_d[m], subs_h[m], (size_t)nnz * sizeof(long long), st));
    }
    ACC_CUDA(cudaMalloc(&vals_d, nnz * sizeof(double)));
    ACC_TRY(skt_upload(vals_d, vals_h, (size_t)nnz * sizeof(double), st));
    ACC_CUDA(cudaMalloc(&subs_dd, ndim * sizeof(long long *)));
    ACC_CUDA(cudaMalloc(&base_d, ndim * sizeof(long long)));
    ACC_CUDA(cudaMalloc(&stride_d, ndim * sizeof(long long)));
    ACC_CUDA(cudaMemcpyAsync(subs_dd, subs_d, ndim * sizeof(long long *), cudaMemcpyHostToDevice, st));
    ACC_CUDA(cudaMemcpyAsync(base_d, base, ndim * sizeof(long long), cudaMemcpyHostToDevice, st));
    ACC_CUDA(cudaMemcpyAsync(stride_d, stride, ndim * sizeof(long long), cudaMemcpyHostToDevice, st));
    ACC_TRY(alloc_sorted(S, nnz));
    accum_key_kernel<<<blocks_for(nnz), 256, 0, st>>>((const long long *const *)subs_dd, base_d, stride_d, ndim, nnz, S.key, S.perm);
    ++g_launches;
    long long nd = 0;
    ACC_TRY(sort_and_number(S, nnz, bits, st, &nd));
    ACC_CUDA(cudaMalloc(&okey_d, nd * sizeof(unsigned long long)));
    ACC_CUDA(cudaMalloc(&oval_d, nd * sizeof(double)));
    run_reduce_kernel<<<blocks_for(nnz), 256, 0, st>>>(S.key_sorted, S.perm_sorted, S.slot, vals_d, nnz, op, okey_d, oval_d);
    ++g_launches;
    ACC_CUDA(cudaGetLastError());
    ACC_CUDA(cudaMemcpyAsync(out_keys_h, okey_d, nd * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    ACC_CUDA(cudaMemcpyAsync(out_vals_h, oval_d, nd * sizeof(double), cudaMemcpyDeviceToHost, st));
    ACC_CUDA(cudaStreamSynchronize(st));
#undef ACC_TRY
#undef ACC_CUDA
    *out_nnz = nd;
    cleanup();
    return SKT_OK;
}

extern "C" int skt_compact_keys(const uint64_t *keys_d, int64_t n, int key_bits, int32_t *ids_d, int64_t *ndistinct,
                                skt_stream stream) {
    SKT_REQUIRE(keys_d && ids_d && ndistinct, SKT_EINVAL, "NULL argument");
    SKT_REQUIRE(n >= 1 && n < (1LL << 31) - 1, SKT_ESHAPE, "n %lld outside [1, 2^31)", (long long)n);
    SKT_REQUIRE(key_bits >= 1 && key_bits <= 64, SKT_EINVAL, "key_bits must be in [1, 64]");
    cudaStream_t st = (cudaStream_t)stream;
    SortedKeys S;
    int rc = alloc_sorted(S, n);
    if (rc == SKT_OK) {
        cudaMemcpyAsync(S.key, keys_d, n * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st);
        iota_u32_kernel<<<blocks_for(n), 256, 0, st>>>(S.perm, n);
        ++g_launches;
        long long nd = 0;
        rc = sort_and_number(S, n, key_bits, st, &nd);
        if (rc == SKT_OK) {
            scatter_ids_kernel<<<blocks_for(n), 256, 0, st>>>(S.key_sorted, S.perm_sorted, S.slot, n, ids_d);
            ++g_launches;
            if (cudaStreamSynchronize(st) != cudaSuccess) rc = SKT_ECUDA;
            *ndistinct = nd;
        }
    }
    S.release();
    return rc;
}

/*
 * sktensor_b200.h -- C ABI of libsktensor_b200.so: the B200 (sm_100a) kernels behind
 * scikit-tensor's CP-ALS / MTTKRP / HOOI hot path.
 *
 * The reference (mnick/scikit-tensor) is pure Python and has no FFI; the boundary this
 * library replaces is the set of NumPy/SciPy calls made underneath the Python methods
 * named below.  Each entry point cites the reference file:line it stands in for.  The
 * reference-side binding (ctypes) is shown in INTEGRATION.md and implemented in
 * scikit-tensor_b200/sktensor/_lib.py.
 *
 * Conventions
 *   - every function returns SKT_OK (0) or a negative SKT_E* code; it never throws or
 *     aborts.  skt_last_error() returns a thread-local, human-readable message.
 *   - pointers named *_d are DEVICE pointers (any allocator: cudaMalloc, torch, ...),
 *     pointers named *_h are HOST pointers.  Unless stated, buffers are borrowed for the
 *     duration of the call only.  All matrices are row-major FP64; tensors are C-order FP64.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Device
 *     entry points only enqueue work; they do not synchronise.  *_host entry points copy
 *     in, compute, copy out and synchronise before returning.
 *   - workspaces: call the matching *_workspace() query, hand in at least that many bytes,
 *     256-byte aligned.  Workspaces carry no state between calls.
 *   - no CPU fallback exists anywhere in this library.
 */
#ifndef SKTENSOR_B200_H
#define SKTENSOR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SKT_API __attribute__((visibility("default")))
#else
#define SKT_API
#endif

#define SKT_MAX_NDIM 8
#define SKT_MAX_RANK 64          /* single-CTA / cluster solve and update kernels                                    */
#define SKT_MAX_WIDE_RANK 512    /* generic solve path (skt_gram / skt_hadamard_pinv / skt_solve_stats /
                                    skt_normalise_gram / skt_cp_fit) assembled from the large-matrix kernels       */

#define SKT_OK            0
#define SKT_EINVAL       -1      /* bad argument (NULL, ndim, mode, rank ...)          */
#define SKT_ESHAPE       -2      /* shape too large / inconsistent                      */
#define SKT_ECUDA        -3      /* CUDA runtime error (message has the cudaError text) */
#define SKT_EWORKSPACE   -4      /* workspace too small                                 */
#define SKT_ENONFINITE   -5      /* NaN/Inf met where the reference raises ValueError   */
#define SKT_EUNSUPPORTED -6

typedef void *skt_stream;

/* ---------------------------------------------------------------- library / device */
SKT_API int         skt_version(void);
SKT_API const char *skt_last_error(void);
/* Number of kernels this library has launched in the calling process (bench.py's gpu_launches). */
SKT_API uint64_t skt_launch_count(void);
/* SM count, compute capability and opt-in shared memory of the current device. */
SKT_API int skt_device_info(int *sm_count, int *cc_major, int *cc_minor, size_t *smem_optin);

/* ---------------------------------------------------------------- tensor ingest
 * Copy a (pageable) host buffer to the device through a pool of pinned staging buffers filled by several
 * host threads, each overlapping its memcpy with its own asynchronous DMA.  The reference starts every
 * call from a host ndarray (dtensor.__new__, sktensor/dtensor.py:48-50); this is the first step of
 * cp_als / tucker_hooi / uttkrp here.  The copy is ordered on `stream`; src_h is not referenced after
 * the call returns.                                                                               */
SKT_API int skt_upload(void *dst_d, const void *src_h, size_t nbytes, skt_stream stream);

/* ---------------------------------------------------------------- dense MTTKRP
 * M[i_n, r] = sum_{all other i} X[i_0..i_{N-1}] * prod_{m != n} U_m[i_m, r]
 * Replaces dtensor.uttkrp (sktensor/dtensor.py:163-167) = unfold (dtensor.py:103-150) +
 * khatrirao (core.py:333-378) + dot.  Neither the unfolding nor the Khatri-Rao matrix is
 * formed.  U_d is a HOST array of `ndim` device pointers; U_d[mode] is ignored and may be
 * NULL (cp.py:194,143).  M_d is (shape[mode] x rank).                                   */
SKT_API size_t skt_mttkrp_dense_workspace(const int64_t *shape, int ndim, int rank, int mode);
SKT_API int skt_mttkrp_dense(const double *X_d, const int64_t *shape, int ndim,
                     const double *const *U_d, int rank, int mode, double *M_d,
                     void *ws_d, size_t ws_bytes, skt_stream stream);
/* Straightforward one-thread-per-output kernel; used by tests as an independent GPU check. */
SKT_API int skt_mttkrp_dense_naive(const double *X_d, const int64_t *shape, int ndim,
                           const double *const *U_d, int rank, int mode, double *M_d,
                           skt_stream stream);
/* Host-buffer form (end-to-end through the ABI): uploads X and the factors, runs
 * skt_mttkrp_dense, downloads M.  X_h is C-order.                                       */
SKT_API int skt_mttkrp_dense_host(const double *X_h, const int64_t *shape, int ndim,
                          const double *const *U_h, int rank, int mode, double *M_h);

/* ---------------------------------------------------------------- dimension-tree second stage
 * The N MTTKRPs of one cp.als sweep (cp.py:142-143) share work: with T the partial contraction of X
 * against the factors of every mode OUTSIDE a group of p consecutive modes (T is
 * dims[0] x ... x dims[p-1] x rank, row-major; produced by skt_mttkrp_dense on a reshaped view of X),
 *   M[i_mode, r] = sum_{other i of the group} T[i_0..i_{p-1}, r] * prod_{m != mode} W_m[i_m, r]
 * equals dtensor.uttkrp(U, mode) (dtensor.py:163-167) for that group mode.  W_d: HOST array of p device
 * pointers (dims[m] x rank), W_d[mode] ignored.  Deterministic (no atomics).                       */
SKT_API size_t skt_kr_reduce_workspace(const int64_t *dims, int p, int rank, int mode);
SKT_API int skt_kr_reduce(const double *T_d, const int64_t *dims, int p, const double *const *W_d, int rank,
                  int mode, double *M_d, void *ws_d, size_t ws_bytes, skt_stream stream);

/* The same second stage with the per-split partial slabs left un-summed in slabs_d ([*nslabs_out][dims[mode]][rank],
 * at most skt_kr_reduce_workspace bytes); skt_cp_update_cluster_slabs adds them in split order while staging M.  */
SKT_API int skt_kr_reduce_slabs(const double *T_d, const int64_t *dims, int p, const double *const *W_d, int rank,
                        int mode, double *slabs_d, size_t slabs_bytes, int *nslabs_out, skt_stream stream);

/* Dense reconstruction of a Kruskal tensor, out[i_0..i_{N-1}] = sum_r lambda_r prod_m U_m[i_m, r] in C order
 * (ktensor.toarray / totensor, sktensor/ktensor.py:152-175: `dot(lmbda, khatrirao(U).T).reshape(shape)`), without
 * forming the Khatri-Rao matrix.  lambda_d may be NULL (ones).                                                  */
SKT_API int skt_ktensor_full(const double *const *U_d, const double *lambda_d, const int64_t *shape, int ndim, int rank,
                     double *out_d, skt_stream stream);

/* ---------------------------------------------------------------- sparse (COO) MTTKRP
 * Replaces sptensor.uttkrp (sktensor/sptensor.py:216-229) and its per-rank-column
 * ttv/_ttv_compute passes (sptensor.py:153-177: gather, unique, searchsorted, bincount).
 * skt_coo_build sorts the non-zeros once per output mode (stable radix sort on the mode's
 * subscript, int32 indices) and keeps the sorted copies on the device; skt_mttkrp_coo then
 * runs a segmented reduction with no global atomics.  Duplicate coordinates are summed
 * (sptensor.py:172); rows without non-zeros come out as zeros (sptensor.py:221).        */
typedef struct skt_coo skt_coo;
/* subs_h: `ndim` host arrays of nnz int64 subscripts; vals_h: nnz doubles.  modes_mask:
 * bit n set = build the sorted copy for output mode n (0 = all modes).                  */
SKT_API int skt_coo_build_host(const int64_t *const *subs_h, const double *vals_h, int64_t nnz,
                       const int64_t *shape, int ndim, unsigned modes_mask, skt_coo **out);
/* Same, from device arrays (int64 subscripts). */
SKT_API int skt_coo_build(const int64_t *const *subs_d, const double *vals_d, int64_t nnz,
                  const int64_t *shape, int ndim, unsigned modes_mask, skt_stream stream,
                  skt_coo **out);
SKT_API int     skt_coo_destroy(skt_coo *t);
SKT_API int64_t skt_coo_nnz(const skt_coo *t);
SKT_API size_t  skt_coo_device_bytes(const skt_coo *t);
SKT_API size_t  skt_mttkrp_coo_workspace(const skt_coo *t, int rank, int mode);
SKT_API int skt_mttkrp_coo(const skt_coo *t, const double *const *U_d, int rank, int mode,
                   double *M_d, void *ws_d, size_t ws_bytes, skt_stream stream);
/* Unsorted-COO kernel with FP64 global atomics; tests use it as an independent GPU check. */
SKT_API int skt_mttkrp_coo_atomic(const skt_coo *t, const double *const *U_d, int rank, int mode,
                          double *M_d, skt_stream stream);
/* <P, X> for sparse X: sum_nnz val * sum_r lambda_r prod_m U_m[i_m, r]
 * (ktensor.innerprod, ktensor.py:128-150 -> sptensor.py:154-163).  out_d: one double.   */
SKT_API size_t skt_coo_innerprod_workspace(const skt_coo *t, int rank);
SKT_API int skt_coo_innerprod(const skt_coo *t, const double *const *U_d, const double *lambda_d,
                      int rank, double *out_d, void *ws_d, size_t ws_bytes, skt_stream stream);
/* sum of squares of the stored values (sptensor.norm, sptensor.py:274-282, squared). */
SKT_API int skt_coo_sumsq(const skt_coo *t, double *out_d, void *ws_d, size_t ws_bytes, skt_stream stream);

/* ---------------------------------------------------------------- reductions
 * out_d[0] = sum_i x[i]^2   (dtensor.norm, dtensor.py:152-161, squared).  Deterministic.  */
SKT_API size_t skt_sumsq_workspace(int64_t n);
SKT_API int skt_sumsq(const double *x_d, int64_t n, double *out_d, void *ws_d, size_t ws_bytes,
              skt_stream stream);

/* out_d[0] = sum_i sum_r lambda_r U[i,r] M[i,r] (lambda_d NULL = ones).  With M the MTTKRP of
 * the last mode this is <P, X> (ktensor.innerprod, ktensor.py:128-150) without the reference's
 * rank-many passes over X.  Workspace: skt_sumsq_workspace(rows*rank).                    */
SKT_API int skt_weighted_dot(const double *U_d, const double *M_d, const double *lambda_d, int64_t rows,
                     int rank, double *out_d, void *ws_d, size_t ws_bytes, skt_stream stream);

/* ---------------------------------------------------------------- CP-ALS normal equations
 * G = U^T U for a tall (rows x rank) factor (cp.py:146, ktensor.py:125).                 */
SKT_API size_t skt_gram_workspace(int64_t rows, int rank);
SKT_API int skt_gram(const double *U_d, int64_t rows, int rank, double *G_d, void *ws_d,
             size_t ws_bytes, skt_stream stream);
/* P = pinv( hadamard_{m != mode} G_m )  with scipy.linalg.pinv's cut-off
 * rtol = rank * eps * sigma_max (cp.py:144-147).  G_d holds ndim (rank x rank) Grams.
 * Computed by a cyclic Jacobi eigen-decomposition in one CTA.  info_d[0] receives the
 * numerical rank kept, or -1 if Y held NaN/Inf (the reference raises ValueError there). */
SKT_API int skt_hadamard_pinv(const double *G_d, int ndim, int rank, int mode, double *P_d,
                      int *info_d, skt_stream stream);
/* Phase A of the factor update: Unew = M P (cp.py:147) and the per-column statistic the
 * normalisation needs: sum of squares when first_iter != 0 (cp.py:149-150), signed maximum
 * otherwise (cp.py:151-152).  colstat_d: rank doubles.                                   */
SKT_API size_t skt_solve_stats_workspace(int64_t rows, int rank);
SKT_API int skt_solve_stats(const double *M_d, const double *P_d, int64_t rows, int rank,
                    int first_iter, double *Unew_d, double *colstat_d, void *ws_d,
                    size_t ws_bytes, skt_stream stream);
/* Phase B: lambda = sqrt(colstat) (first_iter) or max(colstat, 1) (cp.py:150-153);
 * U = Unew / lambda in place (cp.py:154); G = U^T U of the normalised factor; and, when
 * Mip_d != NULL, ip_d[0] = sum_r lambda_r sum_i U[i,r] Mip[i,r]  (= <P,X> when Mip is the
 * MTTKRP of the last mode; replaces ktensor.innerprod, ktensor.py:128-150).  A multi-GPU
 * caller all-reduces colstat between the two phases and G / ip afterwards.              */
SKT_API size_t skt_normalise_gram_workspace(int64_t rows, int rank);
SKT_API int skt_normalise_gram(double *U_d, const double *colstat_d, int64_t rows, int rank,
                       int first_iter, double *lambda_d, double *G_d, const double *Mip_d,
                       double *ip_d, void *ws_d, size_t ws_bytes, skt_stream stream);
/* The whole inter-MTTKRP block of cp.als for a factor that fits in shared memory (cp.py:144-159) in one
 * single-CTA launch: P = pinv(hadamard_{m != mode} G_m), Unew = M P, lambda, U = Unew / lambda (U_d),
 * G[mode] = U^T U (written into G_d), and optionally ip_d[0] = <P,X> (M being the MTTKRP of the last mode,
 * ktensor.py:128-150) and fit_d[0..1] = (fit, ||P||^2) (cp.py:157-159).  skt_cp_update_fused_fits() tells
 * whether (rows, rank) qualifies; tall factors use skt_hadamard_pinv / skt_solve_stats /
 * skt_normalise_gram.  info_d[0]: rank kept by the pseudo-inverse, -1 on a non-finite Gram product.   */
SKT_API int skt_cp_update_fused_fits(int64_t rows, int rank);
SKT_API int skt_cp_update_fused(const double *M_d, int64_t rows, int rank, int ndim, int mode, int first_iter,
                        double *G_d, double *U_d, double *lambda_d, int *info_d, double *ip_d,
                        const double *normX2_d, double *fit_d, skt_stream stream);
/* The same update given P = pinv(Y) (from skt_hadamard_pinv): one 8-CTA thread-block cluster splits the factor
 * rows, exchanges the column statistics and partial Grams through distributed shared memory in CTA-rank order
 * (deterministic), and also writes <P,X> and the fit when asked.  cp.als needs pinv(Y) of mode n only after the
 * MTTKRP of mode n (cp.py:143-147) while Y depends on the OTHER modes' Grams alone, so the caller can run
 * skt_hadamard_pinv on a second stream during that MTTKRP and keep only this short kernel on the critical path. */
SKT_API int skt_cp_update_cluster_fits(int64_t rows, int rank);
SKT_API int skt_cp_update_cluster(const double *M_d, const double *P_d, int64_t rows, int rank, int ndim, int mode,
                          int first_iter, double *G_d, double *U_d, double *lambda_d, double *ip_d,
                          const double *normX2_d, double *fit_d, skt_stream stream);
/* Exchange form of the update for a factor whose rows are spread over ranks (the sharded mode) or too tall for
 * shared memory (sparse tensors).  skt_cp_update_exchange: Unew = M P (cp.py:147, not yet normalised) and one record
 * ex_d = [colstat (R) | Unew^T Unew (R x R) | <Unew, M>] of skt_cp_exchange_record(R) doubles.  The ranks all-gather
 * their records (one collective instead of three all-reduces); skt_cp_update_finish then forms, in rank order,
 * lambda (cp.py:149-153), G[mode] = (sum_w Unew_w^T Unew_w) ./ (lambda lambda^T), <P,X>, the fit (cp.py:157-159),
 * and divides the local rows by lambda -- skipped when every lambda is 1.  world = 1 for an unsharded factor.   */
SKT_API size_t skt_cp_exchange_record(int rank);
SKT_API size_t skt_cp_update_exchange_workspace(int64_t rows, int rank);
SKT_API int skt_cp_update_exchange(const double *M_d, const double *P_d, int64_t rows, int rank, int first_iter,
                           double *Unew_d, double *ex_d, void *ws_d, size_t ws_bytes, skt_stream stream);
SKT_API int skt_cp_update_finish(const double *ex_all_d, int world, int first_iter, int64_t rows, int rank, int ndim,
                         int mode, double *U_d, double *lambda_d, double *G_d, double *ip_d,
                         const double *normX2_d, double *fit_d, skt_stream stream);
/* Fused reduce + update over NVLink peer memory.  Same as skt_cp_update_cluster / skt_cp_update_finish, but the
 * operand that would have come out of an NCCL collective is read straight from the peers' HBM inside the kernel:
 * Mparts_d / ex_parts_d is a DEVICE array of `nparts` pointers, entry w = rank w's buffer mapped into this process
 * (symmetric memory; torch.distributed._symmetric_memory: handle.buffer_ptrs_dev).  The partial MTTKRPs (or exchange
 * records) are added in rank order, so every rank gets bitwise the same factor.  The caller orders the ranks with a
 * cross-rank barrier on the stream (handle.barrier()) before the call.                                          */
SKT_API int skt_cp_update_cluster_parts(const double *const *Mparts_d, int nparts, const double *P_d, int64_t rows,
                                int rank, int ndim, int mode, int first_iter, double *G_d, double *U_d,
                                double *lambda_d, double *ip_d, const double *normX2_d, double *fit_d,
                                skt_stream stream);
SKT_API int skt_cp_update_finish_parts(const double *const *ex_parts_d, int world, int first_iter, int64_t rows, int rank,
                               int ndim, int mode, double *U_d, double *lambda_d, double *G_d, double *ip_d,
                               const double *normX2_d, double *fit_d, skt_stream stream);
/* skt_cp_update_cluster with M given as `nslabs` local partial slabs of rows*rank doubles each (skt_kr_reduce_slabs). */
SKT_API int skt_cp_update_cluster_slabs(const double *slabs_d, int nslabs, const double *P_d, int64_t rows, int rank,
                                int ndim, int mode, int first_iter, double *G_d, double *U_d, double *lambda_d,
                                double *ip_d, const double *normX2_d, double *fit_d, skt_stream stream);
/* ---------------------------------------------------------------- the whole CP-ALS loop in one call
 * cp.als for a dense tensor whose factors all fit the single-launch update kernel (sktensor/cp.py:132-171: per mode
 * uttkrp -> Gram-Hadamard pinv solve -> normalise; per iteration fit + convergence test `itr > 0 and |fit - fitold| <
 * conv`).  U_d: storage for all ndim factors (I_m x rank, row-major); bit m of init_mask set = U_d[m] holds an initial
 * factor (cp._init leaves the first one out, cp.py:190-205; its storage is overwritten before it is read).
 * fit_full = 0 reproduces fit_method=None (fit = iteration index, no early exit before max_iter unless conv > 1).
 * On return: *fit_h, *itr_h (index of the last iteration), exectimes_h[0..*itr_h] wall seconds (may be NULL),
 * lambda_d and the factors on the device.  SKT_ENONFINITE where scipy's pinv raises ValueError (cp.py:147).      */
SKT_API int    skt_cp_als_run_fits(const int64_t *shape, int ndim, int rank);
SKT_API size_t skt_cp_als_run_workspace(const int64_t *shape, int ndim, int rank);
SKT_API int    skt_cp_als_run(const double *X_d, const int64_t *shape, int ndim, int rank, double *const *U_d,
                      unsigned init_mask, int max_iter, double conv, int fit_full, double *lambda_d, void *ws_d,
                      size_t ws_bytes, double *fit_h, int *itr_h, double *exectimes_h, skt_stream stream);

/* ---------------------------------------------------------------- opt-in: sliced contraction on the INT8 tensor cores
 * The two partial contractions of the dense 3-way dimension tree (the unfolded product of dtensor.uttkrp,
 * sktensor/dtensor.py:117-137, on the 2-way views  T[row, r] = sum_c X[.., c, ..] U[c, r]) computed FP64-accurately
 * by error-free slicing: X is cut ONCE into `nslices` (6: 48-bit, 5: 40-bit) planes of balanced byte digits under one
 * power-of-two scale; every pass streams those planes through tcgen05.mma.kind::i8 with INT32 accumulators in tensor
 * memory and recombines them in FP64 (csrc/slice_gemm.cu).  The default path (skt_mttkrp_dense) stays FP64 DMMA.
 * skt_planes_supported: 1 when skt_planes_ttm accepts (shape, mode, rank): mode first or last, rank <= 32, contracted
 * mode <= 512, 16-byte aligned plane rows.  skt_planes_ttm: T_d (rows x rank, rows = the other modes in memory order).
 * Error bound |dT| <= 2^-(8 nslices - 1) (max|X| sum_c |U[c,r]| + max_c |U[c,r]| sum_c |X|); non-finite input -> NaN. */
typedef struct skt_planes skt_planes;
SKT_API int    skt_planes_supported(const int64_t *shape, int ndim, int mode, int rank);
SKT_API int    skt_planes_create(const double *X_d, const int64_t *shape, int ndim, int nslices, skt_planes **out,
                         skt_stream stream);
SKT_API int    skt_planes_destroy(skt_planes *p);
SKT_API size_t skt_planes_device_bytes(const skt_planes *p);
SKT_API int    skt_planes_nslices(const skt_planes *p);
SKT_API size_t skt_planes_ttm_workspace(const skt_planes *p, int mode, int rank);
SKT_API int    skt_planes_ttm(const skt_planes *p, int mode, const double *U_d, int rank, double *T_d, void *ws_d,
                      size_t ws_bytes, skt_stream stream);
/* 3-way tensors, mode = 0: skt_planes_ttm with the second stage of the dimension tree folded into its epilogue.  W_d = the
 * factor of the MIDDLE mode (shape[1] x rank).  Writes slabs of M_2[k,r] = sum_j (sum_i X[i,j,k] U[i,r]) W[j,r]; T_d is not
 * written (may be NULL).  The MTTKRP is the SUM of the slabs (skt_planes_fused_slabs slabs of *slab_rows x rank, contiguous),
 * the form skt_cp_update_cluster_slabs consumes.  Needs shape[2] to be a multiple of 128; skt_planes_fused_slabs returns 0
 * when the shape or the mode is not supported.                                                                        */
SKT_API int    skt_planes_fused_slabs(const skt_planes *p, int mode, int rank, int64_t *slab_rows);
SKT_API int    skt_planes_ttm_fused(const skt_planes *p, int mode, const double *U_d, const double *W_d, int rank, double *T_d,
                            double *slabs_d, size_t slabs_bytes, void *ws_d, size_t ws_bytes, skt_stream stream);
/* General tree pass on the same planes: X seen as the matrix [P x Q], P = prod(shape[:split]), Q = prod(shape[split:]).
 * group 0: T_d[P x rank] = X KR(U_split .. U_{ndim-1});  group 1: T_d[Q x rank] = X^T KR(U_0 .. U_{split-1}), KR = the
 * Khatri-Rao product of the group's factors in X's memory order (the khatrirao + dot of dtensor.uttkrp, dtensor.py:117-137
 * with core.py:333-378, restricted to one group of modes; its rows are formed on the fly, never stored in FP64).  U_d holds
 * ndim factor pointers (only the group's are read).  rank <= 64, Q a multiple of 16; contractions of any length (the
 * factor planes stream through shared memory, INT32 accumulators are flushed every 16384 products).                 */
SKT_API int    skt_planes_pass_supported(const int64_t *shape, int ndim, int split, int rank);
SKT_API size_t skt_planes_pass_workspace(const skt_planes *p, int split, int group, int rank);
SKT_API int    skt_planes_pass(const skt_planes *p, int split, int group, const double *const *U_d, int rank, double *T_d,
                       void *ws_d, size_t ws_bytes, skt_stream stream);

/* ---------------------------------------------------------------- COO de-duplication (accumfun)
 * sptensor(subs, vals, accumfun=...) merges duplicate coordinates (sktensor/sptensor.py:80-84 -> utils.accum,
 * sktensor/utils.py:5-26: lexsort with the LAST mode most significant, `func` over every run).  Host arrays in, host
 * arrays out: the distinct coordinates come back as linear keys  sum_m (sub_m - base_m) * prod_{m' < m} extent_m'
 * in ascending order (the reference's output order) with the reduced values; op 0 = sum, 1 = max, 2 = min, applied
 * in storage order.  out_keys_h / out_vals_h must hold nnz entries.                                              */
SKT_API int skt_coo_accumulate_host(const int64_t *const *subs_h, const double *vals_h, int64_t nnz, const int64_t *base,
                            const int64_t *extent, int ndim, int op, uint64_t *out_keys_h, double *out_vals_h,
                            int64_t *out_nnz);
/* ids_d[i] = rank of keys_d[i] among the distinct keys (ascending); *ndistinct = their number.  Used to number the
 * non-empty columns of a sparse unfolding (sptensor.unfold, sktensor/sptensor.py:197-214) for the sparse nvecs
 * operator.  Synchronises the stream.                                                                         */
SKT_API int skt_compact_keys(const uint64_t *keys_d, int64_t n, int key_bits, int32_t *ids_d, int64_t *ndistinct,
                     skt_stream stream);

/* ---------------------------------------------------------------- symmetric eigensolver (nvecs)
 * The `rank` leading eigenvectors of a symmetric n x n matrix (row-major, both triangles stored), eigenvalues in
 * DESCENDING order, each column sign-fixed so that its largest-magnitude entry is positive when `flipsign` != 0:
 * the arithmetic of sktensor.core.nvecs after the Gram product (sktensor/core.py:284-293: scipy.linalg.eigh with
 * eigvals=(N-rank, N-1), columns reversed, then flipsign, core.py:297-306).  One kernel on one thread-block cluster:
 * Householder tridiagonalisation with the matrix in distributed shared memory, Sturm multisection, inverse iteration,
 * back-transformation.  W_d (k eigenvalues) and info_d (0 ok, -1 non-finite input) may be NULL.  n <= 2048.        */
SKT_API size_t skt_symeig_workspace(int n, int k);
SKT_API int skt_symeig_top(const double *A_d, int n, int k, int flipsign, double *W_d, double *U_d, int *info_d,
                   void *ws_d, size_t ws_bytes, skt_stream stream);

/* Leave `n` SMs out of the persistent grids (dense MTTKRP) so that a small kernel on another stream can run
 * beside them; returns the previous value.  0 restores the full chip.                                      */
SKT_API int skt_set_reserved_sms(int n);
/* fit = 1 - (normX2 + ||P||^2 - 2 ip) / normX2 with ||P||^2 = sum(lambda lambda^T *
 * hadamard_m G_m) (cp.py:157-159, ktensor.py:113-126).  out_d[0] = fit, out_d[1] = ||P||^2. */
SKT_API int skt_cp_fit(const double *G_d, int ndim, int rank, const double *lambda_d,
               const double *ip_d, const double *normX2_d, double *out_d, skt_stream stream);

/* ---------------------------------------------------------------- Tucker / HOOI
 * Y = X x_mode V^T (transp != 0: V is I_mode x p) or X x_mode V (V is p x I_mode)
 * (dtensor._ttm_compute, dtensor.py:58-76), without the transpose/reshape copies.        */
SKT_API int skt_ttm_dense(const double *X_d, const int64_t *shape, int ndim, const double *V_d,
                  int64_t p, int mode, int transp, double *Y_d, skt_stream stream);
/* Y = X x_0 V^T (V is shape[0] x p) with the NEW mode stored LAST: Y has shape (shape[1], ..., shape[ndim-1], p).
 * Same numbers as dtensor._ttm_compute (dtensor.py:58-76) up to that rotation of the mode order; tucker.hooi
 * keeps the rotated order between the steps of its TTM chain (tucker.py:103, core.py:99-102), so the big first
 * contraction runs as an M-major GEMM on the TMA / FP64 tensor-core kernel and no transpose is materialised. */
SKT_API int skt_ttm_dense_first_rot(const double *X_d, const int64_t *shape, int ndim, const double *V_d, int64_t p,
                            double *Y_d, skt_stream stream);
/* G = X_(mode) X_(mode)^T  (I_mode x I_mode), the Gram inside core.nvecs (core.py:279,285). */
SKT_API size_t skt_unfold_gram_workspace(const int64_t *shape, int ndim, int mode);
SKT_API int skt_unfold_gram(const double *X_d, const int64_t *shape, int ndim, int mode, double *G_d,
                    void *ws_d, size_t ws_bytes, skt_stream stream);

/* ---------------------------------------------------------------- measurement helpers
 * Peak probes used by bench.py for the roofline denominators that MEASURED_PEAKS.json
 * does not carry: FP64 DMMA (mma.sync m8n8k4) and FP64 FMA issue rate, in TFLOP/s.       */
SKT_API int skt_probe_fp64_peaks(double *dmma_tflops, double *dfma_tflops);

#ifdef __cplusplus
}
#endif
#endif /* SKTENSOR_B200_H */

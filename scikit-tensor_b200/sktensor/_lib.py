"""ctypes binding of ``libsktensor_b200.so`` (C ABI: ``include/sktensor_b200.h``).

This is the whole FFI surface of the package: every numerical call of the CP-ALS /
MTTKRP / HOOI path goes through one of the functions declared here.  There is no CPU
fallback -- if the library cannot be loaded, or no CUDA device is visible, the first
use raises ``RuntimeError``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get(
    'SKTENSOR_B200_LIB', os.path.join(os.path.dirname(_HERE), 'lib', 'libsktensor_b200.so'))

SKT_MAX_NDIM = 8
SKT_MAX_RANK = 64

_ERRNAMES = {-1: 'SKT_EINVAL', -2: 'SKT_ESHAPE', -3: 'SKT_ECUDA', -4: 'SKT_EWORKSPACE',
             -5: 'SKT_ENONFINITE', -6: 'SKT_EUNSUPPORTED'}

c_i64p = C.POINTER(C.c_int64)
c_voidpp = C.POINTER(C.c_void_p)
c_dp = C.c_void_p          # device / host double* passed as an integer address
c_stream = C.c_void_p

# name -> (restype, argtypes); mirrors include/sktensor_b200.h declaration by declaration
PROTOTYPES = {
    'skt_version': (C.c_int, []),
    'skt_last_error': (C.c_char_p, []),
    'skt_launch_count': (C.c_uint64, []),
    'skt_device_info': (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                  C.POINTER(C.c_size_t)]),
    'skt_upload': (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, c_stream]),
    'skt_mttkrp_dense_workspace': (C.c_size_t, [c_i64p, C.c_int, C.c_int, C.c_int]),
    'skt_mttkrp_dense': (C.c_int, [c_dp, c_i64p, C.c_int, c_voidpp, C.c_int, C.c_int, c_dp,
                                   C.c_void_p, C.c_size_t, c_stream]),
    'skt_mttkrp_dense_naive': (C.c_int, [c_dp, c_i64p, C.c_int, c_voidpp, C.c_int, C.c_int, c_dp, c_stream]),
    'skt_mttkrp_dense_host': (C.c_int, [c_dp, c_i64p, C.c_int, c_voidpp, C.c_int, C.c_int, c_dp]),
    'skt_kr_reduce_workspace': (C.c_size_t, [c_i64p, C.c_int, C.c_int, C.c_int]),
    'skt_kr_reduce': (C.c_int, [c_dp, c_i64p, C.c_int, c_voidpp, C.c_int, C.c_int, c_dp, C.c_void_p, C.c_size_t,
                                c_stream]),
    'skt_kr_reduce_slabs': (C.c_int, [c_dp, c_i64p, C.c_int, c_voidpp, C.c_int, C.c_int, c_dp, C.c_size_t,
                                      C.POINTER(C.c_int), c_stream]),
    'skt_ktensor_full': (C.c_int, [c_voidpp, c_dp, c_i64p, C.c_int, C.c_int, c_dp, c_stream]),
    'skt_coo_build_host': (C.c_int, [c_voidpp, c_dp, C.c_int64, c_i64p, C.c_int, C.c_uint,
                                     C.POINTER(C.c_void_p)]),
    'skt_coo_build': (C.c_int, [c_voidpp, c_dp, C.c_int64, c_i64p, C.c_int, C.c_uint, c_stream,
                                C.POINTER(C.c_void_p)]),
    'skt_coo_destroy': (C.c_int, [C.c_void_p]),
    'skt_coo_nnz': (C.c_int64, [C.c_void_p]),
    'skt_coo_device_bytes': (C.c_size_t, [C.c_void_p]),
    'skt_mttkrp_coo_workspace': (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    'skt_mttkrp_coo': (C.c_int, [C.c_void_p, c_voidpp, C.c_int, C.c_int, c_dp, C.c_void_p, C.c_size_t,
                                 c_stream]),
    'skt_mttkrp_coo_atomic': (C.c_int, [C.c_void_p, c_voidpp, C.c_int, C.c_int, c_dp, c_stream]),
    'skt_coo_innerprod_workspace': (C.c_size_t, [C.c_void_p, C.c_int]),
    'skt_coo_innerprod': (C.c_int, [C.c_void_p, c_voidpp, c_dp, C.c_int, c_dp, C.c_void_p, C.c_size_t,
                                    c_stream]),
    'skt_coo_sumsq': (C.c_int, [C.c_void_p, c_dp, C.c_void_p, C.c_size_t, c_stream]),
    'skt_sumsq_workspace': (C.c_size_t, [C.c_int64]),
    'skt_sumsq': (C.c_int, [c_dp, C.c_int64, c_dp, C.c_void_p, C.c_size_t, c_stream]),
    'skt_weighted_dot': (C.c_int, [c_dp, c_dp, c_dp, C.c_int64, C.c_int, c_dp, C.c_void_p, C.c_size_t, c_stream]),
    'skt_gram_workspace': (C.c_size_t, [C.c_int64, C.c_int]),
    'skt_gram': (C.c_int, [c_dp, C.c_int64, C.c_int, c_dp, C.c_void_p, C.c_size_t, c_stream]),
    'skt_hadamard_pinv': (C.c_int, [c_dp, C.c_int, C.c_int, C.c_int, c_dp, C.c_void_p, c_stream]),
    'skt_solve_stats_workspace': (C.c_size_t, [C.c_int64, C.c_int]),
    'skt_solve_stats': (C.c_int, [c_dp, c_dp, C.c_int64, C.c_int, C.c_int, c_dp, c_dp, C.c_void_p,
                                  C.c_size_t, c_stream]),
    'skt_normalise_gram_workspace': (C.c_size_t, [C.c_int64, C.c_int]),
    'skt_normalise_gram': (C.c_int, [c_dp, c_dp, C.c_int64, C.c_int, C.c_int, c_dp, c_dp, c_dp, c_dp,
                                     C.c_void_p, C.c_size_t, c_stream]),
    'skt_cp_update_fused_fits': (C.c_int, [C.c_int64, C.c_int]),
    'skt_cp_update_fused': (C.c_int, [c_dp, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_dp,
                                      C.c_void_p, c_dp, c_dp, c_dp, c_stream]),
    'skt_cp_update_cluster_fits': (C.c_int, [C.c_int64, C.c_int]),
    'skt_cp_update_cluster': (C.c_int, [c_dp, c_dp, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_dp,
                                        c_dp, c_dp, c_dp, c_stream]),
    'skt_cp_exchange_record': (C.c_size_t, [C.c_int]),
    'skt_cp_update_exchange_workspace': (C.c_size_t, [C.c_int64, C.c_int]),
    'skt_cp_update_exchange': (C.c_int, [c_dp, c_dp, C.c_int64, C.c_int, C.c_int, c_dp, c_dp, C.c_void_p, C.c_size_t,
                                         c_stream]),
    'skt_cp_update_finish': (C.c_int, [c_dp, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_dp,
                                       c_dp, c_dp, c_dp, c_stream]),
    'skt_cp_update_cluster_parts': (C.c_int, [C.c_void_p, C.c_int, c_dp, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, c_dp,
                                              c_dp, c_dp, c_dp, c_dp, c_dp, c_stream]),
    'skt_cp_update_finish_parts': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int, c_dp, c_dp,
                                             c_dp, c_dp, c_dp, c_dp, c_stream]),
    'skt_cp_update_cluster_slabs': (C.c_int, [c_dp, C.c_int, c_dp, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, c_dp,
                                              c_dp, c_dp, c_dp, c_dp, c_dp, c_stream]),
    'skt_set_reserved_sms': (C.c_int, [C.c_int]),
    'skt_cp_fit': (C.c_int, [c_dp, C.c_int, C.c_int, c_dp, c_dp, c_dp, c_dp, c_stream]),
    'skt_ttm_dense': (C.c_int, [c_dp, c_i64p, C.c_int, c_dp, C.c_int64, C.c_int, C.c_int, c_dp, c_stream]),
    'skt_ttm_dense_first_rot': (C.c_int, [c_dp, c_i64p, C.c_int, c_dp, C.c_int64, c_dp, c_stream]),
    'skt_unfold_gram_workspace': (C.c_size_t, [c_i64p, C.c_int, C.c_int]),
    'skt_unfold_gram': (C.c_int, [c_dp, c_i64p, C.c_int, C.c_int, c_dp, C.c_void_p, C.c_size_t, c_stream]),
    'skt_probe_fp64_peaks': (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    'skt_planes_supported': (C.c_int, [c_i64p, C.c_int, C.c_int, C.c_int]),
    'skt_planes_create': (C.c_int, [c_dp, c_i64p, C.c_int, C.c_int, C.POINTER(C.c_void_p), c_stream]),
    'skt_planes_destroy': (C.c_int, [C.c_void_p]),
    'skt_planes_device_bytes': (C.c_size_t, [C.c_void_p]),
    'skt_planes_nslices': (C.c_int, [C.c_void_p]),
    'skt_planes_ttm_workspace': (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    'skt_planes_ttm': (C.c_int, [C.c_void_p, C.c_int, c_dp, C.c_int, c_dp, C.c_void_p, C.c_size_t, c_stream]),
    'skt_planes_fused_slabs': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int64)]),
    'skt_planes_ttm_fused': (C.c_int, [C.c_void_p, C.c_int, c_dp, c_dp, C.c_int, c_dp, c_dp, C.c_size_t, C.c_void_p, C.c_size_t,
                                       c_stream]),
    'skt_planes_pass_supported': (C.c_int, [c_i64p, C.c_int, C.c_int, C.c_int]),
    'skt_planes_pass_workspace': (C.c_size_t, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    'skt_planes_pass': (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_voidpp, C.c_int, c_dp, C.c_void_p, C.c_size_t, c_stream]),
    'skt_coo_accumulate_host': (C.c_int, [c_voidpp, c_dp, C.c_int64, c_i64p, c_i64p, C.c_int, C.c_int, C.c_void_p, c_dp,
                                          C.POINTER(C.c_int64)]),
    'skt_compact_keys': (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.POINTER(C.c_int64), c_stream]),
    'skt_cp_als_run_fits': (C.c_int, [c_i64p, C.c_int, C.c_int]),
    'skt_cp_als_run_workspace': (C.c_size_t, [c_i64p, C.c_int, C.c_int]),
    'skt_cp_als_run': (C.c_int, [c_dp, c_i64p, C.c_int, C.c_int, c_voidpp, C.c_uint, C.c_int, C.c_double, C.c_int, c_dp,
                                 C.c_void_p, C.c_size_t, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p, c_stream]),
    'skt_symeig_workspace': (C.c_size_t, [C.c_int, C.c_int]),
    'skt_symeig_top': (C.c_int, [c_dp, C.c_int, C.c_int, C.c_int, c_dp, c_dp, C.c_void_p, C.c_void_p, C.c_size_t, c_stream]),
}

_lib = None


class SktError(RuntimeError):
    """A non-zero status from libsktensor_b200 (message from ``skt_last_error``)."""

    def __init__(self, code, msg):
        self.code = code
        RuntimeError.__init__(self, '%s (%d): %s' % (_ERRNAMES.get(code, 'SKT_E?'), code, msg))


def load():
    """Load the shared library once and attach the prototypes.  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            'libsktensor_b200.so not found at %s -- build it with `make -C scikit-tensor_b200` '
            '(or `python __graft_entry__.py build`); this package has no CPU fallback' % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)        # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    """Map a status code to the reference's exception vocabulary (ValueError for bad input)."""
    if rc == 0:
        return
    msg = load().skt_last_error().decode('utf-8', 'replace')
    if rc in (-1, -2, -5):
        raise ValueError('%s: %s' % (_ERRNAMES[rc], msg))
    raise SktError(rc, msg)


def shape_array(shape):
    return (C.c_int64 * len(shape))(*[int(s) for s in shape])


def ptr_array(ptrs):
    """Host array of ``void*`` from integer addresses (``None``/0 -> NULL)."""
    return (C.c_void_p * len(ptrs))(*[C.c_void_p(int(p)) if p else None for p in ptrs])

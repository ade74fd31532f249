"""ctypes binding of libcaesar_b200.so -- the same C ABI (include/caesar_b200.h) that the Julia shim reaches with
`ccall`.  No numerics happen in this file; there is no CPU fallback: if the shared library is missing or the
machine has no B200 the calls raise."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# CB2_LIB selects an experiment build of the same library (bench/micro/build_variant.py); the product is the in-tree default
SO_PATH = os.environ.get("CB2_LIB") or os.path.join(_HERE, "libcaesar_b200.so")
HEADER_PATH = os.path.join(_HERE, "..", "include", "caesar_b200.h")

CB2_FP32, CB2_FP64 = 0, 1
CB2_MAX_DIM, CB2_MAX_DENS, CB2_MAX_PRODUCT_N = 8, 16, 16384


class CB2Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libcaesar_b200 error {code}: {msg}")
        self.code = code


class Density(C.Structure):
    _fields_ = [("N", C.c_int32), ("pts", C.c_void_p), ("bw", C.c_void_p), ("w", C.c_void_p)]


class ProductDesc(C.Structure):
    _fields_ = [("dens", C.POINTER(Density)), ("D", C.c_int32), ("Nout", C.c_int32), ("niter", C.c_int32),
                ("stream", C.c_uint32), ("sample_offset", C.c_int32), ("pts_out", C.c_void_p),
                ("labels_out", C.c_void_p), ("bw_out", C.c_void_p)]


class ConvDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reverse", C.c_int32), ("N", C.c_int32), ("n_src", C.c_int32), ("src", C.c_void_p),
                ("aux", C.c_void_p), ("n_aux", C.c_int32), ("stream", C.c_uint32), ("mean", C.c_double * 6),
                ("chol", C.c_double * 36), ("out", C.c_void_p)]


_lib = None


def lib():
    """Loads the in-tree shared library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(f"{SO_PATH} is missing: run `python caesar.jl_b200/build.py` (nvcc, sm_100a). "
                              "There is no CPU fallback.")
        L = C.CDLL(SO_PATH)
        vp, i32, f64, u64, u32 = C.c_void_p, C.c_int, C.c_double, C.c_uint64, C.c_uint32
        sig = {
            "cb2_create": (i32, [i32, i32, C.POINTER(vp)]),
            "cb2_destroy": (None, [vp]),
            "cb2_last_error": (C.c_char_p, [vp]),
            "cb2_precision": (i32, [vp]),
            "cb2_set_stream": (i32, [vp, vp]),
            "cb2_synchronize": (i32, [vp]),
            "cb2_kernel_launches": (u64, [vp]),
            "cb2_device_sm_count": (i32, [vp]),
            "cb2_last_stage_ms": (f64, [vp, C.c_char_p]),
            "cb2_mmd": (i32, [vp, vp, i32, vp, i32, i32, f64, vp]),
            "cb2_mmd_manifold": (i32, [vp, vp, i32, vp, i32, i32, vp, f64, f64, vp]),
            "cb2_mmd_sums": (i32, [vp, vp, i32, vp, i32, i32, f64, i32, i32, i32, i32, vp]),
            "cb2_mmd_sums_dev": (i32, [vp, vp, i32, vp, i32, i32, f64, i32, i32, i32, i32, vp]),
            "cb2_mmd_sums_shard": (i32, [vp, vp, i32, vp, i32, i32, f64, i32, i32, vp]),
            "cb2_mmd_sums_shard_dev": (i32, [vp, vp, i32, vp, i32, i32, f64, i32, i32, vp]),
            "cb2_mmd_se_batch": (i32, [vp, vp, i32, vp, i32, i32, f64, vp, i32, i32, vp, vp]),
            "cb2_mmd_se_batch_dev": (i32, [vp, vp, i32, vp, i32, i32, f64, vp, i32, i32, vp, vp]),
            "cb2_kde_eval": (i32, [vp, vp, vp, vp, i32, i32, vp, i32, i32, vp]),
            "cb2_kde_bw_loo": (i32, [vp, vp, i32, i32, vp, vp]),
            "cb2_kde_bw_loo_batch": (i32, [vp, vp, vp, i32, i32, vp, vp]),
            "cb2_kde_sample": (i32, [vp, vp, vp, vp, i32, i32, vp, i32, u64, u32, vp]),
            "cb2_product": (i32, [vp, C.POINTER(Density), i32, vp, i32, i32, i32, u64, u32, i32, vp, vp]),
            "cb2_product_batch": (i32, [vp, C.POINTER(ProductDesc), i32, vp, i32, u64]),
            "cb2_product_batch_dev": (i32, [vp, C.POINTER(ProductDesc), i32, vp, i32, u64]),
            "cb2_product_batch_submit": (i32, [vp, C.POINTER(ProductDesc), i32, vp, i32, u64, C.POINTER(C.c_int64)]),
            "cb2_wait": (i32, [vp, C.c_int64]),
            "cb2_tree_debug": (i32, [vp, C.POINTER(Density), i32, i32, vp, vp, vp, vp, vp, vp]),
            "cb2_approx_conv_batch": (i32, [vp, C.POINTER(ConvDesc), i32, u64]),
            "cb2_se_residual_batch": (i32, [vp, i32, vp, vp, vp, i32, vp]),
            "cb2_debug_counter": (i32, [vp, i32, i32, C.POINTER(u64)]),
            "cb2_set_level_down": (i32, [vp, i32]),
            "cb2_device_alloc": (i32, [vp, C.c_size_t, C.POINTER(vp)]),
            "cb2_device_free": (i32, [vp, vp]),
            "cb2_upload": (i32, [vp, vp, vp, C.c_size_t]),
            "cb2_download": (i32, [vp, vp, vp, C.c_size_t]),
            "cb2_approx_conv_batch_dev": (i32, [vp, C.POINTER(ConvDesc), i32, u64]),
            "cb2_kde_bw_loo_batch_dev": (i32, [vp, vp, vp, i32, i32, vp, vp]),
            "cb2_icp_align_batch": (i32, [vp, vp, i32, vp, i32, vp, i32, i32, i32, f64, f64, i32, vp, vp]),
            "cb2_scatter_align": (i32, [vp, C.POINTER(Density), C.POINTER(Density), i32, i32, f64, vp, i32, u64, i32, f64,
                                        vp, vp, vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


EXPORTS = ["cb2_create", "cb2_destroy", "cb2_last_error", "cb2_precision", "cb2_set_stream", "cb2_synchronize",
           "cb2_kernel_launches", "cb2_device_sm_count", "cb2_last_stage_ms", "cb2_mmd", "cb2_mmd_sums",
           "cb2_mmd_sums_dev", "cb2_mmd_se_batch", "cb2_mmd_se_batch_dev", "cb2_kde_eval", "cb2_kde_bw_loo",
           "cb2_kde_bw_loo_batch", "cb2_kde_sample", "cb2_product", "cb2_product_batch", "cb2_product_batch_dev",
           "cb2_product_batch_submit", "cb2_wait", "cb2_tree_debug", "cb2_scatter_align", "cb2_approx_conv_batch",
           "cb2_mmd_manifold", "cb2_device_alloc", "cb2_device_free", "cb2_upload", "cb2_download", "cb2_approx_conv_batch_dev",
           "cb2_kde_bw_loo_batch_dev", "cb2_icp_align_batch", "cb2_se_residual_batch",
           "cb2_mmd_sums_shard", "cb2_mmd_sums_shard_dev", "cb2_debug_counter", "cb2_set_level_down"]


def f64(x):
    return np.ascontiguousarray(x, dtype=np.float64)


def ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a))  # raw address (device pointer / torch data_ptr)


def circ_arr(circular, d):
    if circular is None:
        return None
    c = np.ascontiguousarray(circular, dtype=np.uint8)
    if c.size != d:
        raise ValueError(f"circular mask has {c.size} entries for d={d}")
    return c

"""ctypes binding of libhdn.so (include/hdn.h) and the in-tree build recipe.

The CUDA library is the product: if it is missing this module raises -- there is no CPU
fallback anywhere in the package.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
# HDN_LIB: load another build of the same sources (A/B variants of a kernel timed side by side, scripts/build_variants.sh)
LIB_PATH = os.environ.get("HDN_LIB") or os.path.join(_HERE, "libhdn.so")
SOURCES = ["api.cu", "conv_simt.cu", "conv_tc.cu", "conv_tc_wgrad.cu", "conv_tc2_wgrad.cu", "elementwise.cu", "postproc.cu", "augment.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into h-denseunet_b200/libhdn.so (in-tree)."""
    srcs = [os.path.join(_CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(_CSRC, "hdn_common.cuh"), os.path.join(_CSRC, "tc_common.cuh"),
                   os.path.join(_HERE, "..", "include", "hdn.h")]
    deps = [d for d in deps if os.path.exists(d)]
    if not force and os.path.exists(LIB_PATH):
        t = os.path.getmtime(LIB_PATH)
        if all(os.path.getmtime(d) <= t for d in deps):
            return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("HDN_NVCC_EXTRA", "").split()      # e.g. -DHDN_TC_TIMING (per-role cycle counters)
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-o", LIB_PATH] + srcs + ["-lcuda"]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True, cwd=_CSRC)
    return LIB_PATH


# ----------------------------------------------------------------------------- structs
class Tensor(C.Structure):
    _fields_ = [("p", C.c_void_p), ("ldc", C.c_int), ("coff", C.c_int)]


class Src(C.Structure):
    _fields_ = [("t", Tensor), ("D", C.c_int), ("H", C.c_int), ("W", C.c_int),
                ("ud", C.c_int), ("uh", C.c_int), ("uw", C.c_int),
                ("pa", C.c_void_p), ("pb", C.c_void_p), ("relu", C.c_int)]


class Conv(C.Structure):
    _fields_ = [("N", C.c_int), ("D", C.c_int), ("H", C.c_int), ("W", C.c_int),
                ("Cin", C.c_int), ("Cout", C.c_int),
                ("kd", C.c_int), ("kh", C.c_int), ("kw", C.c_int),
                ("sd", C.c_int), ("sh", C.c_int), ("sw", C.c_int),
                ("pd", C.c_int), ("ph", C.c_int), ("pw", C.c_int),
                ("nsrc", C.c_int), ("src", Src * 2),
                ("w", C.c_void_p), ("bias", C.c_void_p), ("y", Tensor),
                ("stat_sum", C.c_void_p), ("stat_sq", C.c_void_p),
                ("drop_keep", C.c_float), ("drop_seed", C.c_uint64), ("precision", C.c_int),
                ("ws", C.c_void_p), ("ws_bytes", C.c_int64)]


class DgradEpi(C.Structure):
    _fields_ = [("dx", Tensor), ("du", C.c_void_p), ("mode", C.c_int), ("accumulate", C.c_int),
                ("s1", C.c_void_p), ("s2", C.c_void_p), ("center", C.c_void_p)]


class Pool(C.Structure):
    _fields_ = [("kind", C.c_int), ("N", C.c_int), ("D", C.c_int), ("H", C.c_int), ("W", C.c_int),
                ("C", C.c_int), ("pool_d", C.c_int), ("src", Src), ("y", Tensor), ("argidx", C.c_void_p)]


class BnFold(C.Structure):
    _fields_ = [("C", C.c_int), ("mode", C.c_int), ("count", C.c_double),
                ("sum", C.c_void_p), ("sumsq", C.c_void_p),
                ("mov_mean", C.c_void_p), ("mov_var", C.c_void_p),
                ("gamma", C.c_void_p), ("beta", C.c_void_p),
                ("sgamma", C.c_void_p), ("sbeta", C.c_void_p),
                ("eps", C.c_float), ("momentum", C.c_float),
                ("a", C.c_void_p), ("b", C.c_void_p), ("mean", C.c_void_p), ("rstd", C.c_void_p)]


class BnGrad(C.Structure):
    _fields_ = [("C", C.c_int), ("mode", C.c_int), ("count", C.c_double),
                ("s1", C.c_void_p), ("s2", C.c_void_p), ("mean", C.c_void_p), ("rstd", C.c_void_p),
                ("gamma", C.c_void_p), ("beta", C.c_void_p), ("sgamma", C.c_void_p),
                ("dgamma", C.c_void_p), ("dbeta", C.c_void_p), ("dsgamma", C.c_void_p),
                ("dsbeta", C.c_void_p), ("k0", C.c_void_p), ("k1", C.c_void_p), ("k2", C.c_void_p)]


class Aug(C.Structure):
    _fields_ = [("vol", C.c_void_p), ("seg", C.c_void_p), ("vol_i16", C.c_int32), ("VS", C.c_int32), ("VH", C.c_int32), ("VW", C.c_int32),
                ("a0", C.c_int32), ("b0", C.c_int32), ("c0", C.c_int32), ("ch", C.c_int32), ("cw", C.c_int32), ("cs", C.c_int32),
                ("m00", C.c_int32), ("m01", C.c_int32), ("m10", C.c_int32), ("m11", C.c_int32), ("o0", C.c_int32), ("o1", C.c_int32),
                ("mean", C.c_float), ("out_h", C.c_int32), ("out_w", C.c_int32), ("xs_s", C.c_int64), ("xs_h", C.c_int64), ("xs_w", C.c_int64),
                ("ys0", C.c_int32), ("yns", C.c_int32)]


EXPORTS = [
    "hdn_last_error", "hdn_version", "hdn_launch_count", "hdn_conv_fprop", "hdn_conv_dgrad", "hdn_conv_wgrad",
    "hdn_conv_tc_supported", "hdn_conv_tc_workspace", "hdn_conv_tc_plan", "hdn_pool_fwd", "hdn_pool_bwd", "hdn_bn_fold", "hdn_bn_param_grad",
    "hdn_bn_bwd_apply", "hdn_dropout_bwd", "hdn_col_stats", "hdn_wce_accum", "hdn_wce_grad",
    "hdn_triplets", "hdn_cat4", "hdn_cat4_bwd", "hdn_sgd_nesterov", "hdn_dp_reduce_sgd",
    "hdn_window_accumulate", "hdn_window_finalize", "hdn_dev_malloc", "hdn_dev_free",
    "hdn_ipc_get_handle", "hdn_ipc_open", "hdn_ipc_close", "hdn_set_switch", "hdn_dp_signal", "hdn_dp_wait", "hdn_layout_nhws_to_nshw",
    "hdn_post_threshold", "hdn_post_dilate", "hdn_post_largest_component", "hdn_post_fill_holes", "hdn_post_and", "hdn_post_compose", "hdn_aug_sample",
]

_lib = None


class HdnError(RuntimeError):
    pass


def load():
    """Load libhdn.so; raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HdnError("libhdn.so is not built (%s missing). Run __graft_entry__.build(); "
                       "there is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.hdn_last_error.restype = C.c_char_p
    lib.hdn_launch_count.restype = C.c_longlong
    vp, i32, i64, f32, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint64
    sig = {
        "hdn_conv_fprop": [C.POINTER(Conv), vp],
        "hdn_conv_dgrad": [C.POINTER(Conv), C.POINTER(DgradEpi), vp],
        "hdn_conv_wgrad": [C.POINTER(Conv), vp, vp, vp],
        "hdn_conv_tc_supported": [C.POINTER(Conv), i32],
        "hdn_conv_tc_plan": [C.POINTER(Conv), i32, C.POINTER(C.c_int32)],
        "hdn_pool_fwd": [C.POINTER(Pool), vp],
        "hdn_pool_bwd": [C.POINTER(Pool), C.POINTER(DgradEpi), vp],
        "hdn_bn_fold": [C.POINTER(BnFold), vp],
        "hdn_bn_param_grad": [C.POINTER(BnGrad), vp],
        "hdn_bn_bwd_apply": [vp, Tensor, Tensor, i64, i32, vp, vp, vp, vp, i32, vp],
        "hdn_dropout_bwd": [Tensor, i64, i32, f32, u64, vp],
        "hdn_col_stats": [Tensor, i64, i32, vp, vp, vp],
        "hdn_wce_accum": [vp, vp, i64, i32, i64, i32, i32, vp, vp],
        "hdn_wce_grad": [vp, vp, vp, i64, i32, i64, i32, i32, vp, f32, vp],
        "hdn_triplets": [vp, vp, i32, i32, i64, i32, vp],
        "hdn_cat4": [vp, vp, vp, i64, f32, vp],
        "hdn_cat4_bwd": [vp, vp, i64, f32, i32, vp],
        "hdn_sgd_nesterov": [vp, vp, vp, i64, f32, f32, f32, vp],
        "hdn_dp_reduce_sgd": [vp, vp, vp, i32, i32, i64, i64, f32, f32, f32, vp],
        "hdn_window_accumulate": [vp, vp, vp, i32, i64, i32, vp],
        "hdn_window_finalize": [vp, vp, i32, i64, vp],
        "hdn_dev_malloc": [C.POINTER(vp), i64],
        "hdn_dev_free": [vp],
        "hdn_ipc_get_handle": [vp, vp],
        "hdn_ipc_open": [vp, C.POINTER(vp)],
        "hdn_ipc_close": [vp],
        "hdn_set_switch": [C.c_char_p, i32],
        "hdn_dp_signal": [vp, i32, i32, C.c_uint32, vp],
        "hdn_dp_wait": [vp, i32, C.c_uint32, vp],
        "hdn_layout_nhws_to_nshw": [vp, vp, i32, i32, i32, i32, i32, vp],
        "hdn_post_threshold": [vp, vp, vp, vp, i64, f32, f32, vp],
        "hdn_post_dilate": [vp, vp, i32, i32, i32, vp],
        "hdn_post_largest_component": [vp, vp, i32, i32, i32, vp, i64, vp],
        "hdn_post_fill_holes": [vp, vp, i32, i32, i32, vp, i64, vp],
        "hdn_post_and": [vp, vp, vp, i64, vp],
        "hdn_post_compose": [vp, vp, vp, i64, vp],
        "hdn_aug_sample": [C.POINTER(Aug), vp, vp, vp, vp, vp],
    }
    for name, args in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.hdn_conv_tc_workspace.argtypes = [C.POINTER(Conv), i32]
    lib.hdn_conv_tc_workspace.restype = C.c_int64
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = _lib.hdn_last_error().decode() if _lib is not None else "?"
        raise HdnError("%s failed (%d): %s" % (what or "hdn call", rc, msg))

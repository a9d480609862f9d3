"""ctypes binding of libcamvid_b200.so (the C ABI declared in include/camvid_b200.h).

The library is the product: there is no CPU or PyTorch fallback behind these calls. If the shared object is missing
or an entry point fails, a RuntimeError is raised with the library's own message.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcamvid_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "camvid_b200.h")


class View(ctypes.Structure):
    """cvb_view: NHWC bf16 view, element strides."""
    _fields_ = [("ptr", ctypes.c_void_p), ("n", ctypes.c_int32), ("h", ctypes.c_int32), ("w", ctypes.c_int32),
                ("c", ctypes.c_int32), ("sn", ctypes.c_int64), ("sh", ctypes.c_int64), ("sw", ctypes.c_int64)]


class ConvEpilogue(ctypes.Structure):
    """cvb_conv_epilogue."""
    _fields_ = [("stat_partials", ctypes.c_void_p), ("scale", ctypes.c_void_p), ("shift", ctypes.c_void_p),
                ("relu", ctypes.c_int32), ("bwd_y", View), ("bwd_scale", ctypes.c_void_p),
                ("bwd_shift", ctypes.c_void_p), ("bwd_partials", ctypes.c_void_p)]


class Comm(ctypes.Structure):
    """cvb_comm: peer pointers of the symmetric gradient buffer / flag pad."""
    _fields_ = [("peer_bufs_host", ctypes.POINTER(ctypes.c_void_p)), ("peer_flags_host", ctypes.POINTER(ctypes.c_void_p)),
                ("multicast_buf", ctypes.c_void_p), ("rank", ctypes.c_int32), ("world", ctypes.c_int32)]


_P = ctypes.c_void_p
_I = ctypes.c_int
_L = ctypes.c_int64
_F = ctypes.c_float

# name -> (restype, argtypes); must list every CVB_API symbol of include/camvid_b200.h (tests check this)
SIGNATURES = {
    "cvb_last_error": (ctypes.c_char_p, []),
    "cvb_abi_version": (_I, []),
    "cvb_sm_count": (_I, []),
    "cvb_shutdown": (None, []),
    "cvb_nchw_f32_to_nhwc_bf16": (_I, [_P, _I, View, _P]),
    "cvb_nhwc_bf16_to_nchw_f32": (_I, [View, _P, _I, _P]),
    "cvb_im2col3x3_nchw_f32": (_I, [_P, _I, View, _P]),
    "cvb_pack_weights_fprop": (_I, [_P, _I, _I, _I, _I, _I, _P, _P]),
    "cvb_pack_weights_dgrad": (_I, [_P, _I, _I, _I, _I, _P, _P]),
    "cvb_pack_weights_batch": (_I, [_P, _I, _I, _I, _P]),
    "cvb_conv_stat_rows": (_I, []),
    "cvb_conv3x3_fprop": (_I, [View, _P, _I, View, ctypes.POINTER(ConvEpilogue), _P]),
    "cvb_conv3x3_fprop_fuses_bwd_stats": (_I, [View, View, _I]),
    "cvb_adamw_chunk_elems": (_I, []),
    "cvb_adamw_step": (_I, [_P, _P, _I, _F, _F, _F, _F, _F, _L, _P]),
    "cvb_adamw_factors": (_I, [_F, _F, _F, _F, _F, _L, _P]),
    "cvb_adamw_step_dev": (_I, [_P, _P, _I, _P, _P]),
    "cvb_conv3x3_wgrad_workspace_bytes": (_L, [View, View, _I]),
    "cvb_conv3x3_wgrad": (_I, [View, View, _I, _P, _I, _I, _P, _L, _P]),
    "cvb_bn_stats": (_I, [View, _P, _I, _P]),
    "cvb_bn_finalize": (_I, [_P, _I, _I, _I, _L, _P, _P, _P, _P, _P, _F, _F, _P, _P, _P, _P, _P]),
    "cvb_bn_relu_apply": (_I, [View, _P, _P, View, _I, _P]),
    "cvb_bn_relu_bwd_reduce": (_I, [View, View, _P, _P, _P, _I, _I, _P]),
    "cvb_bn_relu_apply_nchw_f32": (_I, [View, _P, _P, _P, _I, _P]),
    "cvb_nchw_f32_to_nhwc_bf16_bn_reduce": (_I, [_P, _I, View, View, _P, _P, _P, _I, _P]),
    "cvb_bn_bwd_finalize": (_I, [_P, _I, _I, _I, _L, _P, _P, _P, _P, _P, _P, _P]),
    "cvb_bn_relu_bwd_apply": (_I, [View, View, _P, _P, _P, View, _I, _P]),
    "cvb_maxpool2x2_fwd": (_I, [View, View, _P, _P]),
    "cvb_bn_relu_maxpool2x2_fwd": (_I, [View, _P, _P, View, View, _P, _P]),
    "cvb_maxpool2x2_bwd": (_I, [View, _P, View, View, _I, _P]),
    "cvb_maxpool2x2_bwd_bn_reduce": (_I, [View, _P, View, _P, _P, View, _I, _P, _I, _P]),
    "cvb_maxunpool2x2_fwd": (_I, [View, _P, View, _P]),
    "cvb_maxunpool2x2_bwd": (_I, [View, _P, View, _P]),
    "cvb_pool_code_to_index": (_I, [_P, _I, _I, _I, _I, _I, _P, _P]),
    "cvb_bilinear2x_fwd": (_I, [View, View, _P]),
    "cvb_bn_relu_bilinear2x_fwd": (_I, [View, _P, _P, View, _P]),
    "cvb_bilinear2x_bwd": (_I, [View, View, _P]),
    "cvb_softmax_ce_nchw_f32": (_I, [_P, _P, _I, _I, _I, _I, _I, _L, _I, _P, _P, _P, _F, _P]),
    "cvb_softmax_ce_nhwc_bf16": (_I, [View, _I, _P, _I, _L, _I, _P, _P, View, _F, _P]),
    "cvb_confusion_matrix": (_I, [_P, _P, _I, _L, _I, _L, _I, _P, _P]),
    "cvb_argmax_confusion_nchw_f32": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _P]),
    "cvb_argmax_confusion_nhwc_bf16": (_I, [View, _I, _P, _I, _P, _P, _P]),
    "cvb_input_stage_u8": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "cvb_comm_flag_words": (_I, []),
    "cvb_allreduce_mean_f32": (_I, [_P, _L, _L, _I, ctypes.c_uint32, _I, _P]),
    "cvb_allreduce_wait": (_I, [_P, _I, ctypes.c_uint32, _P]),
    "cvb_zero_view": (_I, [View, _P]),
}

_lib = None
ABI_VERSION = 4  # CVB_ABI_VERSION of include/camvid_b200.h this binding was written against


def build(verbose=False):
    """Compile libcamvid_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-j8", "-C", CSRC_DIR], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libcamvid_b200.so failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout)
    return LIB_PATH


def load():
    """Load the shared library and attach argument types. Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"(or `make -C {CSRC_DIR}`). camvid_b200 has no fallback path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means header / library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.cvb_abi_version() != ABI_VERSION:
        raise RuntimeError("libcamvid_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().cvb_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")

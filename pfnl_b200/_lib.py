"""ctypes binding of libpfnl_b200.so (include/pfnl_b200.h).

There is no Python/CPU fallback: if the shared library is missing or fails to load, import
of this module raises.  Build it with `python -m pfnl_b200.build` (or __graft_entry__.build()).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# PFNL_B200_LIB: another build of the same library (A/B measurements of kernel variants, tools/ab_build.sh)
LIB_PATH = os.environ.get("PFNL_B200_LIB") or os.path.join(HERE, "libpfnl_b200.so")

NUM_FRAMES = 7
SCALE = 4
NUM_BLOCK = 20
MF = 64
NL_CH = 84

PFNL_OK = 0
ERR_BAD_ARG = -1
ERR_BAD_SHAPE = -2
ERR_CUDA = -3
ERR_UNSUPPORTED_ARCH = -4
ERR_UNIMPLEMENTED = -5
ERR_NO_MEMORY = -6

PREC_FP32 = 0
PREC_TC_FP16X3 = 1
PREC_TC_FP16 = 2
PREC_TC_FP16X3_NLTC = 3
PRECISIONS = {"fp32": PREC_FP32, "fp16x3": PREC_TC_FP16X3, "fp16": PREC_TC_FP16, "fp16x3_nltc": PREC_TC_FP16X3_NLTC}

_fp = C.POINTER(C.c_float)


class PfnlWeights(C.Structure):
    """struct pfnl_weights (host pointers, TF layouts)."""
    _fields_ = [
        ("nl_g_kernel", _fp), ("nl_g_bias", _fp), ("nl_w_kernel", _fp), ("nl_w_bias", _fp),
        ("conv0_kernel", _fp), ("conv0_bias", _fp),
        ("conv1_kernel", _fp * NUM_BLOCK), ("conv1_bias", _fp * NUM_BLOCK),
        ("conv10_kernel", _fp * NUM_BLOCK), ("conv10_bias", _fp * NUM_BLOCK),
        ("conv2_kernel", _fp * NUM_BLOCK), ("conv2_bias", _fp * NUM_BLOCK),
        ("merge1_kernel", _fp), ("merge1_bias", _fp), ("merge2_kernel", _fp), ("merge2_bias", _fp),
    ]


# every symbol include/pfnl_b200.h declares: name -> (restype, argtypes)
_VP = C.c_void_p
_I = C.c_int
SYMBOLS = {
    "pfnl_version": (_I, []),
    "pfnl_last_error": (C.c_char_p, []),
    "pfnl_device_supported": (_I, [_I]),
    "pfnl_create": (_I, [C.POINTER(_VP), _I, C.POINTER(PfnlWeights), _I]),
    "pfnl_destroy": (_I, [_VP]),
    "pfnl_workspace_bytes": (C.c_size_t, [_I, _I, _I, _I]),
    "pfnl_reserve": (_I, [_VP, _I, _I, _I]),
    "pfnl_set_graphs": (_I, [_VP, _I]),
    "pfnl_set_flow": (_I, [_VP, _I]),
    "pfnl_debug_fault": (_I, [C.POINTER(C.c_int)]),
    "pfnl_debug_progress": (_I, [C.POINTER(C.c_int), _I]),
    "pfnl_debug_flow_split": (_I, [_I, _I, C.POINTER(C.c_int)]),
    "pfnl_forward": (_I, [_VP, _VP, _I, _I, _I, _VP, _VP]),
    "pfnl_forward_host": (_I, [_VP, _VP, _I, _I, _I, _VP, _VP]),
    "pfnl_forward_host_submit": (_I, [_VP, _VP, _I, _I, _I, _I, _VP, _VP, C.POINTER(_I)]),
    "pfnl_forward_host_wait": (_I, [_VP, _I]),
    "pfnl_graph_stats": (C.c_longlong, [_VP, _I]),
    "pfnl_mse": (_I, [_VP, _VP, _VP, _I, _I, _I, _VP, _VP]),
    "pfnl_launch_count": (C.c_longlong, [_VP]),
    "pfnl_profile": (_I, [_VP, _I]),
    "pfnl_profile_read": (_I, [_VP, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "pfnl_pack_tokens": (_I, [_VP, _VP, _I, _I, _I, _VP, _VP]),
    "pfnl_nonlocal": (_I, [_VP, _VP, _I, _I, _VP, _VP]),
    "pfnl_depth_to_space": (_I, [_VP, _VP, _I, _I, _I, _I, _I, _VP, _VP]),
    "pfnl_space_to_depth": (_I, [_VP, _VP, _I, _I, _I, _I, _I, _VP, _VP]),
    "pfnl_conv2d_nhwc": (_I, [_VP, _VP, _I, _I, _I, _I, _VP, _VP, _I, _I, _I, _VP, _VP, _VP]),
    "pfnl_bicubic4": (_I, [_VP, _VP, _I, _I, _I, _I, _VP, _VP]),
    "pfnl_pfrb": (_I, [_VP, _I, _VP, _I, _I, _I, _VP, _VP]),
    "pfnl_conv0": (_I, [_VP, _VP, _I, _I, _I, _VP, _VP]),
    "pfnl_convmerge1": (_I, [_VP, _VP, _I, _I, _I, _VP, _VP]),
    "pfnl_downsample4": (_I, [_VP, _VP, _I, _I, _I, _VP, _VP, _VP]),
    "pfnl_gather_windows": (_I, [_VP, _VP, _I, _I, _I, _I, _I, _VP, _VP]),
    "pfnl_quantize_u8": (_I, [_VP, _VP, C.c_longlong, _VP, _VP]),
    "pfnl_msy": (_I, [_VP, _VP, _VP, _I, _I, _I, C.c_float, C.c_float, _I, _I, _VP, _VP]),
    "pfnl_ssim_y": (_I, [_VP, _VP, _VP, _I, _I, _I, C.c_float, C.c_float, _VP, _VP]),
    "pfnl_crc32c": (C.c_uint32, [_VP, C.c_size_t, C.c_uint32]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA library is the product and there is no fallback; "
            "build it with `python -m pfnl_b200.build`")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


class PfnlError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libpfnl_b200 error {code}: {msg}")
        self.code = code


def check(rc):
    if rc != PFNL_OK:
        raise PfnlError(rc, lib.pfnl_last_error().decode("utf-8", "replace"))
    return rc

"""ctypes binding of librfk.so (include/rfk.h).

The library is the product: if it is missing, or a call fails, this module raises — there is no
CPU or PyTorch fallback anywhere in the package.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# RFK_LIB_PATH: developer override used by tools/ to time experimental builds of the same library
LIB_PATH = os.environ.get("RFK_LIB_PATH") or os.path.join(_HERE, "librfk.so")

RFK_F32, RFK_BF16, RFK_F16 = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_ELU = 0, 1, 2
EPI_STD, EPI_BLOCKLN32 = 0, 1

i64 = C.c_int64
i32 = C.c_int32
f32 = C.c_float
vp = C.c_void_p


class RfkAddr(C.Structure):
    _fields_ = [("zs", i64 * 3), ("ms", i64 * 2), ("ns", i64 * 2)]


class RfkGemmDesc(C.Structure):
    _fields_ = [
        ("a", vp), ("b", vp),
        ("ab_dtype", i32), ("act", i32),
        ("M", i64), ("N", i64), ("K", i64),
        ("Z", i64 * 3),
        ("lda", i64), ("ldb", i64),
        ("a_zs", i64 * 3), ("b_zs", i64 * 3),
        ("bias", vp), ("bias_zs", i64 * 3),
        ("alpha", f32), ("epi", i32),
        ("MR", i64), ("NR", i64),
        ("c", vp), ("r0", vp), ("r1", vp),
        ("c_dtype", i32), ("r0_dtype", i32), ("r1_dtype", i32),
        ("ln_eps", f32),
        ("c_addr", RfkAddr), ("r0_addr", RfkAddr), ("r1_addr", RfkAddr),
        ("ln_gamma", vp), ("ln_beta", vp),
    ]


class RfkFavorDesc(C.Structure):
    _fields_ = [
        ("q", vp), ("k", vp), ("v", vp), ("out", vp), ("proj", vp),
        ("io_dtype", i32), ("kind", i32), ("m_features", i32), ("heads", i32),
        ("tokens", i64),
        ("G", i64 * 2), ("gs", i64 * 2), ("ts", i64),
        ("out_gs", i64 * 2), ("out_ts", i64),
    ]


# name -> (restype, argtypes); every symbol include/rfk.h declares
SYMBOLS = {
    "rfk_strerror": (C.c_char_p, [C.c_int]),
    "rfk_version": (C.c_int, []),
    "rfk_launch_count": (C.c_uint64, []),
    "rfk_gemm": (C.c_int, [C.POINTER(RfkGemmDesc), vp]),
    "rfk_layernorm": (C.c_int, [vp, C.c_int, i64, vp, vp, f32, vp, C.c_int, i64, i64, C.c_int, vp]),
    "rfk_layernorm_residual": (C.c_int, [vp, C.c_int, i64, vp, vp, f32, vp, i64, vp, C.c_int, i64, i64, C.c_int, vp]),
    "rfk_dist_mask_logits": (C.c_int, [vp, i64, vp, C.c_int, vp, i64, C.c_int, C.c_int, vp]),
    "rfk_softmax_rows": (C.c_int, [vp, i64, vp, C.c_int, i64, i64, C.c_int, vp]),
    "rfk_tied_att_symmetrize": (C.c_int, [vp, C.c_int, i64, vp, vp, i64, C.c_int, C.c_int, C.c_int, vp]),
    "rfk_poswise_weight": (C.c_int, [vp, i64, vp, i64, C.c_int, f32, vp, vp, i64, f32, vp, C.c_int,
                                     C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "rfk_poswise_weight_stats": (C.c_int, [vp, i64, vp, i64, C.c_int, f32, vp, vp, i64, f32, vp, C.c_int, vp,
                                           C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "rfk_opm_prep": (C.c_int, [vp, vp, vp, vp, C.c_int, i64, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "rfk_pair2att_logits": (C.c_int, [vp, vp, vp, f32, vp, i64, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "rfk_pair2att_logits_rows": (C.c_int, [vp, vp, vp, vp, f32, vp, i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           vp]),
    "rfk_channel_stats": (C.c_int, [vp, C.c_int, vp, C.c_int, i64, C.c_int, vp]),
    "rfk_instnorm_apply": (C.c_int, [vp, C.c_int, vp, vp, vp, f32, vp, C.c_int, C.c_int, vp, C.c_int,
                                     C.c_int, i64, C.c_int, vp]),
    "rfk_favor_attention": (C.c_int, [C.POINTER(RfkFavorDesc), vp]),
    "rfk_conv3x3_nhwc": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "rfk_conv3x3_nhwc_hw": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "rfk_conv3x3_nhwc_f32": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "rfk_conv3x3_nhwc_dil": (C.c_int, [vp, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "rfk_conv3x3_nhwc_f32_dil": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "rfk_pair_symmetrize": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "rfk_msa_embed": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "rfk_pair_embed": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp]),
    "rfk_convert_rows": (C.c_int, [vp, C.c_int, i64, vp, C.c_int, i64, i64, C.c_int, vp]),
}

_lib = None


def load():
    """Load librfk.so (once). Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"librfk.so not found at {LIB_PATH}: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (or rosettafold-pytorch_b200/csrc/build.sh). There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code: int, what: str):
    if code != 0:
        msg = load().rfk_strerror(code).decode()
        raise RuntimeError(f"{what} failed: {msg} (rfk error {code})")


def launch_count() -> int:
    return int(load().rfk_launch_count())

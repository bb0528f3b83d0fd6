"""Error conventions of the C ABI, exercised without a GPU: every entry point must return a code (never crash, never
throw) for null pointers, non-positive sizes and bad dtype codes - those checks come before anything touches the
device - and `rfk_strerror` must name every code. On a GPU-less host a well-formed call must fail loudly with a
non-zero code too (there is no CPU fallback behind the boundary)."""
import ctypes as C

import pytest
import torch

from rosettafold_pytorch_b200 import _lib
from rosettafold_pytorch_b200._lib import RfkFavorDesc, RfkGemmDesc

OK, BAD_DIMS, MISALIGNED, BAD_ARCH, BAD_DTYPE, NULL_PTR = 0, 1, 2, 3, 4, 5
F32, BF16 = 0, 1


@pytest.fixture(scope="module")
def lib():
    return _lib.load()


def test_strerror_names_every_code(lib):
    want = {0: b"ok", 1: b"bad dimensions", 4: b"bad dtype code", 5: b"null pointer", 8: b"unsupported argument combination"}
    for code, text in want.items():
        assert lib.rfk_strerror(code) == text
    for code in range(0, 9):
        assert lib.rfk_strerror(code) not in (None, b"", b"unknown rfk error")
    assert lib.rfk_strerror(77) == b"unknown rfk error"
    assert lib.rfk_strerror(1000 + 2)  # a mapped cudaError_t has a CUDA runtime message


def test_dtype_codes_match_the_binding():
    assert (_lib.RFK_F32, _lib.RFK_BF16) == (F32, BF16)


def test_null_pointers_are_reported_not_dereferenced(lib):
    buf = torch.zeros(64 * 64, dtype=torch.float32)
    p = C.c_void_p(buf.data_ptr())
    assert lib.rfk_gemm(None, None) == NULL_PTR
    d = RfkGemmDesc()
    assert lib.rfk_gemm(C.byref(d), None) == NULL_PTR                      # a, b, c unset
    assert lib.rfk_favor_attention(None, None) == NULL_PTR
    assert lib.rfk_favor_attention(C.byref(RfkFavorDesc()), None) == NULL_PTR
    assert lib.rfk_layernorm(None, F32, 64, None, None, 1e-5, p, F32, 64, 4, 64, None) == NULL_PTR
    assert lib.rfk_layernorm(p, F32, 64, p, None, 1e-5, p, F32, 64, 4, 64, None) == NULL_PTR   # gamma without beta
    assert lib.rfk_softmax_rows(None, 64, p, F32, 64, 4, 64, None) == NULL_PTR
    assert lib.rfk_convert_rows(p, F32, 64, None, BF16, 64, 4, 64, None) == NULL_PTR
    assert lib.rfk_channel_stats(p, F32, None, 1, 16, 64, None) == NULL_PTR
    assert lib.rfk_pair2att_logits(p, None, p, 1e-5, p, 8, 1, 8, 64, 4, None) == NULL_PTR
    assert lib.rfk_pair2att_logits_rows(p, None, p, p, 1e-5, p, 8, 1, 4, 8, 64, 4, None) == NULL_PTR
    assert lib.rfk_poswise_weight(None, 64, p, 64, F32, 1.0, p, None, 0, 1.0, None, F32, 1, 2, 4, 2, 32, None) == NULL_PTR
    assert lib.rfk_poswise_weight_stats(p, 64, p, 64, F32, 1.0, p, None, 0, 1.0, p, F32, None, 1, 2, 4, 2, 32,
                                        None) == NULL_PTR                  # qt without q
    assert lib.rfk_conv3x3_nhwc(None, p, p, BF16, 1, 8, 64, 64, None) == NULL_PTR


def test_bad_sizes_and_dtypes_are_rejected(lib):
    buf = torch.zeros(64 * 64, dtype=torch.float32)
    p = C.c_void_p(buf.data_ptr())
    assert lib.rfk_layernorm(p, F32, 64, None, None, 1e-5, p, F32, 64, -1, 64, None) == BAD_DIMS
    assert lib.rfk_layernorm(p, F32, 64, None, None, 1e-5, p, F32, 64, 4, 0, None) == BAD_DIMS
    assert lib.rfk_layernorm(p, 7, 64, None, None, 1e-5, p, F32, 64, 4, 64, None) == BAD_DTYPE
    assert lib.rfk_layernorm(p, F32, 64, None, None, 1e-5, p, F32, 64, 0, 64, None) == OK      # empty input: nothing to do
    assert lib.rfk_convert_rows(p, F32, 64, p, BF16, 64, 0, 64, None) == OK
    assert lib.rfk_convert_rows(p, F32, 64, p, BF16, 64, 4, 0, None) == BAD_DIMS
    assert lib.rfk_channel_stats(p, F32, p, 0, 16, 64, None) == BAD_DIMS
    assert lib.rfk_pair2att_logits(p, p, p, 1e-5, p, 4, 1, 8, 64, 4, None) == BAD_DIMS           # ld_logits < L
    assert lib.rfk_pair2att_logits(p, p, p, 1e-5, p, 8, 1, 8, 1024, 4, None) == BAD_DIMS         # D > 512
    assert lib.rfk_pair2att_logits_rows(p, p, p, p, 1e-5, p, 8, 1, 9, 8, 64, 4, None) == BAD_DIMS   # more rows than L
    assert lib.rfk_conv3x3_nhwc(p, p, p, BF16, 1, 8, 63, 64, None) == BAD_DIMS                   # C % 8
    assert lib.rfk_conv3x3_nhwc(p, p, p, 9, 1, 8, 64, 64, None) == BAD_DTYPE
    d = RfkGemmDesc()
    d.a = d.b = d.c = buf.data_ptr()
    d.M, d.N, d.K = 64, 64, 0
    assert lib.rfk_gemm(C.byref(d), None) == BAD_DIMS
    d.K = 64
    d.M = 1 << 40
    assert lib.rfk_gemm(C.byref(d), None) == BAD_DIMS                                            # 32-bit tile arithmetic


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a host WITHOUT a GPU")
def test_well_formed_calls_fail_loudly_without_a_gpu(lib):
    """No CPU fallback: a valid call on a GPU-less host returns an error code (architecture check or a mapped CUDA
    runtime error), it neither succeeds nor crashes. The buffers are never touched on the host."""
    buf = torch.zeros(64 * 64, dtype=torch.float32)
    p = C.c_void_p(buf.data_ptr())
    before = lib.rfk_launch_count()
    codes = [lib.rfk_layernorm(p, F32, 64, None, None, 1e-5, p, F32, 64, 4, 64, None),
             lib.rfk_softmax_rows(p, 64, p, F32, 64, 4, 64, None),
             lib.rfk_convert_rows(p, F32, 64, p, BF16, 64, 4, 64, None)]
    d = RfkGemmDesc()
    d.a = d.b = d.c = buf.data_ptr()
    d.ab_dtype, d.c_dtype = BF16, BF16
    d.M, d.N, d.K = 64, 64, 64
    d.Z[0:3] = (1, 1, 1)
    d.lda = d.ldb = 64
    d.MR, d.NR = 64, 64
    d.alpha = 1.0
    d.c_addr.ms[0], d.c_addr.ns[0] = 64, 1
    codes.append(lib.rfk_gemm(C.byref(d), None))
    assert all(c != OK for c in codes), codes
    for c in codes:
        assert lib.rfk_strerror(c)
    assert torch.count_nonzero(buf) == 0
    assert lib.rfk_launch_count() >= before


def test_heads_entry_points_reject_bad_arguments(lib):
    """The entry points added for the prediction heads: dilated convolutions and the pair symmetrisation."""
    buf = torch.zeros(64 * 64, dtype=torch.float32)
    p = C.c_void_p(buf.data_ptr())
    q = C.c_void_p(buf.data_ptr() + 4096)
    F16 = 2
    assert lib.rfk_conv3x3_nhwc_dil(None, BF16, p, p, F32, 1, 8, 8, 64, 64, 2, None) == NULL_PTR
    assert lib.rfk_conv3x3_nhwc_dil(p, BF16, p, p, F32, 1, 8, 8, 64, 64, 0, None) == BAD_DIMS        # dilation < 1
    assert lib.rfk_conv3x3_nhwc_dil(p, BF16, p, p, F32, 1, 8, 8, 64, 64, 65, None) == BAD_DIMS       # dilation > 64
    assert lib.rfk_conv3x3_nhwc_dil(p, F32, p, p, F32, 1, 8, 8, 64, 64, 1, None) == BAD_DTYPE        # image must be 16-bit
    assert lib.rfk_conv3x3_nhwc_dil(p, F16, p, p, 9, 1, 8, 8, 64, 64, 1, None) == BAD_DTYPE
    assert lib.rfk_conv3x3_nhwc_dil(p, BF16, p, p, F32, 1, 8, 8, 60, 64, 1, None) == BAD_DIMS        # C % 8
    assert lib.rfk_conv3x3_nhwc_f32_dil(p, None, p, 1, 8, 8, 64, 64, 1, None) == NULL_PTR
    assert lib.rfk_conv3x3_nhwc_f32_dil(p, p, p, 1, 8, 8, 64, 64, 9, None) == BAD_DIMS              # SIMT form: dilation <= 8
    assert lib.rfk_conv3x3_nhwc_f32_dil(p, p, p, 1, 0, 8, 64, 64, 1, None) == BAD_DIMS
    assert lib.rfk_pair_symmetrize(None, q, F32, 1, 4, 8, None) == NULL_PTR
    assert lib.rfk_pair_symmetrize(p, p, F32, 1, 4, 8, None) == 8                                  # in place: unsupported
    assert lib.rfk_pair_symmetrize(p, q, F32, 1, 4, 6, None) == BAD_DIMS                           # C % 4
    assert lib.rfk_pair_symmetrize(p, q, BF16, 1, 4, 12, None) == BAD_DIMS                         # 16-bit: C % 8
    assert lib.rfk_pair_symmetrize(p, q, 7, 1, 4, 8, None) == BAD_DTYPE
    assert lib.rfk_pair_symmetrize(p, q, F32, 0, 4, 8, None) == BAD_DIMS

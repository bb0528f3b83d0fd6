"""The host side of the C ABI without a GPU: every `rfk_gemm_desc` the CUDA backend would hand to librfk during
a block forward (captured through a stand-in library object) is compared byte for byte with a plain, field-by-field
restatement of include/rfk.h's rules (Z[0] = fastest batch level, zero strides for size-1 / broadcast dims, the
7-D [Z2,Z1,Z0,M1,MR,N1,NR] view split). Guards the hot host path (it runs ~100 times per block) against
optimisations that change what the kernels see."""
import ctypes as C

import rosettafold_pytorch_b200 as rf
from oracle.ops_ref import RefBackend
from rosettafold_pytorch_b200 import ops
from rosettafold_pytorch_b200._lib import RfkGemmDesc
from rosettafold_pytorch_b200.ops import _dt
from tests.helpers import build_block, build_coord_module, load_golden


def _plain_descriptor(a, b, c_view, bias, act, alpha, r0, r1, epi, ln_gamma, ln_beta, ln_eps):
    d = RfkGemmDesc()
    d.a, d.b, d.ab_dtype, d.act = a.data_ptr(), b.data_ptr(), _dt(a), act
    d.M, d.N, d.K = a.shape[3], b.shape[3], a.shape[4]
    for i in range(3):
        d.Z[i] = a.shape[2 - i]
        d.a_zs[i] = a.stride(2 - i) if a.shape[2 - i] > 1 else 0
        d.b_zs[i] = b.stride(2 - i) if b.shape[2 - i] > 1 else 0
    d.lda, d.ldb = a.stride(3), b.stride(3)
    d.bias = None if bias is None else bias.data_ptr()
    d.alpha, d.epi, d.ln_eps = alpha, epi, ln_eps
    d.MR, d.NR = c_view.shape[4], c_view.shape[6]
    for name, t in (("c", c_view), ("r0", r0), ("r1", r1)):
        if t is None:
            continue
        setattr(d, name, t.data_ptr())
        setattr(d, name + "_dtype", _dt(t))
        addr = getattr(d, name + "_addr")
        for i in range(3):
            addr.zs[i] = t.stride(2 - i) if t.shape[2 - i] > 1 else 0
        addr.ms[0], addr.ms[1] = (t.stride(4) if t.shape[4] > 1 else 0), (t.stride(3) if t.shape[3] > 1 else 0)
        addr.ns[0], addr.ns[1] = (t.stride(6) if t.shape[6] > 1 else 0), (t.stride(5) if t.shape[5] > 1 else 0)
    d.ln_gamma = None if ln_gamma is None else ln_gamma.data_ptr()
    d.ln_beta = None if ln_beta is None else ln_beta.data_ptr()
    return C.string_at(C.addressof(d), C.sizeof(d))


def test_gemm_descriptors_of_a_block_forward_match_the_plain_restatement():
    captured = []

    class StandInLib:
        def rfk_gemm(self, dref, stream):
            captured.append(C.string_at(C.addressof(dref._obj), C.sizeof(dref._obj)))
            return 0

    cuda = ops._CudaBackend.__new__(ops._CudaBackend)   # no library load, no GPU
    cuda.lib = StandInLib()
    cuda._stream = lambda t: None
    count = {"n": 0}

    class Checking(RefBackend):
        def gemm(self, *args):
            cuda.gemm(*args)
            assert captured[-1] == _plain_descriptor(*args)
            count["n"] += 1
            return super().gemm(*args)

    prev = ops._set_backend_for_tests(Checking())
    try:
        for mode in ("bf16", "fp32"):
            rf.set_mode(mode)
            blk, _, msa, pair = build_block(dict(d_msa=96, d_pair=72, n_layers=2, B=2, N=5, L=20, seed=3))
            blk(msa, pair)
            mod, _, xyz, state, msa = build_coord_module(load_golden("msa_pair_coord"))
            mod(xyz, state, msa)
    finally:
        ops._set_backend_for_tests(prev)
        rf.set_mode("bf16")
    assert count["n"] > 100


def test_ctypes_signatures_match_the_header():
    """Every prototype of include/rfk.h against the ctypes binding: same number of parameters, pointers bound as
    pointers, 64-bit integers as 64-bit, floats as floats (a drift here corrupts arguments silently)."""
    import os
    import re

    from rosettafold_pytorch_b200 import _lib

    hdr = open(os.path.join(os.path.dirname(_lib.LIB_PATH), "..", "include", "rfk.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = re.findall(r"\b(?:int|const char\*|uint64_t|unsigned long long)\s+(rfk_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", hdr)
    assert {n for n, _ in protos} == set(_lib.SYMBOLS)
    for name, params in protos:
        params = params.strip()
        plist = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
        _, argtypes = _lib.SYMBOLS[name]
        assert len(plist) == len(argtypes), (name, plist, argtypes)
        for p, t in zip(plist, argtypes):
            if "*" in p or "rfk_stream_t" in p:
                assert t is C.c_void_p or issubclass(t, C._Pointer), (name, p, t)
            elif p.startswith("int64_t"):
                assert t is C.c_int64, (name, p, t)
            elif p.startswith("float"):
                assert t is C.c_float, (name, p, t)
            elif p.startswith("int"):
                assert t is C.c_int, (name, p, t)
            else:
                raise AssertionError(f"{name}: unhandled parameter type '{p}'")


def test_ctypes_struct_layouts_match_a_c_compiler(tmp_path):
    """sizeof / offsetof of rfk_gemm_desc, rfk_favor_desc and rfk_addr as gcc lays them out from include/rfk.h vs the
    ctypes Structures of the binding (same field names on both sides)."""
    import os
    import shutil
    import subprocess

    import pytest

    from rosettafold_pytorch_b200 import _lib

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    inc = os.path.abspath(os.path.join(os.path.dirname(_lib.LIB_PATH), "..", "include"))
    structs = {"rfk_gemm_desc": _lib.RfkGemmDesc, "rfk_favor_desc": _lib.RfkFavorDesc, "rfk_addr": _lib.RfkAddr}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "rfk.h"', 'int main(void) {']
    for cname, ct in structs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-I", inc, str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    seen = 0
    for line in out.splitlines():
        cname, field, value = line.split()
        ct = structs[cname]
        want = C.sizeof(ct) if field == "size" else getattr(ct, field).offset
        assert int(value) == want, (cname, field, value, want)
        seen += 1
    assert seen == sum(len(ct._fields_) + 1 for ct in structs.values())

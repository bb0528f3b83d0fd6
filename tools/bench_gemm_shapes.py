"""Per-shape timing of the GEMMs one trunk block issues at the metric config (1,128,512).
Run on the GPU box: python tools/bench_gemm_shapes.py [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rosettafold_pytorch_b200 as rf
from rosettafold_pytorch_b200 import ops
from rosettafold_pytorch_b200.ops import cview

dev = torch.device("cuda:0")
bf, f32 = torch.bfloat16, torch.float32
B, N, L, D, P, H = 1, 128, 512, 384, 288, 12
T, TP = B * N * L, B * L * L
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5


def rnd(*shape, dtype=bf):
    return (torch.randn(*shape, device=dev) * 0.1).to(dtype)


def timeit(name, fn, flops):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print(f"{name:46s} {ms*1e3:9.1f} us  {flops/ms/1e9:8.1f} TFLOP/s")


def lin(name, Tn, Nn, Kn, out_dtype, bias=True, act=0, res=False):
    x, w = rnd(Tn, Kn), rnd(Nn, Kn)
    out = torch.empty(Tn, Nn, dtype=out_dtype, device=dev)
    bvec = torch.randn(Nn, device=dev) if bias else None
    r = torch.randn(Tn, Nn, device=dev) if res else None
    timeit(name, lambda: ops.gemm(x, w, cview(out), bias=bvec, act=act, r0=None if r is None else cview(r)), 2.0 * Tn * Nn * Kn)


print("== MSA side (T=65536)")
lin("tied [q|pk] 384->768 bf16", T, 768, 384, bf)
lin("plain 384->384 bf16", T, 384, 384, bf)
x = rnd(B, N * L, D); w = rnd(D, D); bias = torch.randn(D, device=dev)
kt = torch.empty(B, H, L, N * 32, dtype=bf, device=dev)
timeit("K -> b h l (n d) scatter", lambda: ops.gemm(x, w[None], kt.view(B, H, L, N, 32).permute(0, 3, 2, 1, 4)[None, None], bias=bias), 2.0 * T * D * D)
vt = torch.empty(B, H, N * 32, L, dtype=bf, device=dev)
timeit("V -> b h (n d) l scatter (2-byte stores)", lambda: ops.gemm(x, w[None], vt.view(B, H, N, 32, L).permute(0, 2, 4, 1, 3)[None, None], bias=bias), 2.0 * T * D * D)
qt = rnd(B, H, L, N * 32); logits = torch.empty(B, H, L, L, device=dev)
timeit("tied logits Z=12 512x512x4096", lambda: ops.gemm(qt, kt, logits.view(1, B, H, 1, L, 1, L)), 2.0 * H * L * L * N * 32)
A = rnd(B, H, L, L); o = torch.empty(B, N, L, D, dtype=bf, device=dev)
timeit("tied PV Z=12 512x4096x512 scatter", lambda: ops.gemm(A, vt, o.view(B, N, L, H, 32).permute(0, 3, 2, 1, 4).unsqueeze(2)[None]), 2.0 * H * L * N * 32 * L)
lin("to_out 384->384 +res f32", T, 384, 384, f32, res=True)
lin("FF1 384->1536 relu bf16", T, 1536, 384, bf, act=1)
lin("FF2 1536->384 +res f32", T, 384, 1536, f32, res=True)
lin("performer qkv 384->2304 bf16", T, 2304, 384, bf, bias=False)
lin("performer out 768->384 +res f32", T, 384, 768, f32, res=True)
A4 = rnd(B, 4, L, L); vt4 = rnd(B, 4, N * 96, L); y = torch.empty(B, N, L, D, device=dev); msa = torch.randn(B, N, L, D, device=dev)
asout = lambda t: t.view(B, N, L, 4, 96).permute(0, 3, 2, 1, 4).unsqueeze(2)[None]
timeit("pair->MSA apply Z=4 512x12288x512 +res", lambda: ops.gemm(A4, vt4, asout(y), r0=asout(msa)), 2.0 * 4 * L * N * 96 * L)
print("== pair side (T=262144)")
lin("pair qkv 288->1536 bf16", TP, 1536, 288, bf, bias=False)
lin("pair out 512->288 +res f32", TP, 288, 512, f32, res=True)
lin("pair FF1 288->1152 relu bf16", TP, 1152, 288, bf, act=1)
lin("pair FF2 1152->288 +res f32", TP, 288, 1152, f32, res=True)
lin("OPM linear 1024->288 f32", TP, 288, 1024, f32)
lin("Linear716 (588)->288 f32", TP, 288, 592, f32)
xt, yt = rnd(B, L * 32, N), rnd(B, L * 32, N)
g, bt = torch.ones(1024, device=dev), torch.zeros(1024, device=dev)
oo = torch.empty(B, L, L, 1024, dtype=bf, device=dev)
timeit("OPM outer product + LN1024 (blockln32)", lambda: ops.gemm(xt, yt, oo.view(B, L, L, 32, 32).permute(0, 1, 3, 2, 4)[None, None], epi=ops.EPI_BLOCKLN32), 2.0 * (L * 32) ** 2 * N)

"""Timing of the non-attention pair-side kernels at the metric config (L = 512, d_pair = 288): 3x3 conv,
InstanceNorm statistics / apply, pair2att, row conversion, and the long-K residual GEMMs.
Run on the GPU box: python tools/bench_pair_misc.py [reps]  (A/B env toggles: RFK_CONV_BN, RFK_GEMM_EPI4_ANYK)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rosettafold_pytorch_b200 import ops
from rosettafold_pytorch_b200.ops import cview

dev = torch.device("cuda:0")
bf, f32 = torch.bfloat16, torch.float32
L, P = 512, 288
TP = L * L
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10


def timeit(name, fn, gb=None, gflop=None):
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
    extra = f"{gb / ms:8.0f} GB/s" if gb else f"{gflop / ms:8.1f} TFLOP/s"
    print(f"{name:52s} {ms * 1e3:8.1f} us {extra}")


x = (torch.randn(1, L, L, P, device=dev) * 0.5).to(bf)
w = ops.pack_conv3x3_weight(torch.randn(P, P, 3, 3, device=dev) * 0.02)
y = torch.empty(1, L, L, P, dtype=bf, device=dev)
timeit("conv3x3 288->288 bf16", lambda: ops.conv3x3(x, w, y), gflop=2.0 * TP * P * P * 9 / 1e9)
st = torch.zeros(1, 2, P, dtype=torch.float64, device=dev)
x3 = x.view(1, TP, P)
timeit("channel_stats bf16", lambda: ops.channel_stats(x3, st), gb=TP * P * 2 / 1e6)
g, b_ = torch.randn(P, device=dev), torch.randn(P, device=dev)
st.zero_(); ops.channel_stats(x3, st)
o16 = torch.empty(1, TP, P, dtype=bf, device=dev)
timeit("instnorm_apply bf16 -> bf16, ELU", lambda: ops.instnorm_apply(x3, st, g, b_, 1e-6, o16, elu=True), gb=TP * P * 4 / 1e6)
res = torch.randn(1, TP, P, device=dev)
o32 = torch.empty(1, TP, P, dtype=f32, device=dev)
timeit("instnorm_apply bf16 + res f32 -> f32, ELU", lambda: ops.instnorm_apply(x3, st, g, b_, 1e-6, o32, res=res, elu=True), gb=TP * P * 10 / 1e6)
pair = torch.randn(1, L, L, P, device=dev)
Wf, bfv = torch.randn(4, P, device=dev) * 0.1, torch.randn(4, device=dev)
lg = torch.empty(1, 4, L, L, device=dev)
timeit("pair2att logits", lambda: ops.pair2att_logits(pair, Wf, bfv, 1e-5, lg), gb=TP * P * 4 / 1e6)
h16 = torch.empty(TP, P, dtype=bf, device=dev)
timeit("convert_rows f32 -> bf16", lambda: ops.convert_rows(pair.view(TP, P), h16), gb=TP * P * 6 / 1e6)
for name, K in (("pair FF2 1152->288 +res f32", 1152), ("OPM linear 1024->288 +res f32", 1024), ("pair out 512->288 +res f32", 512)):
    a, wt = (torch.randn(TP, K, device=dev) * 0.1).to(bf), (torch.randn(P, K, device=dev) * 0.1).to(bf)
    out, r, bias = torch.empty(TP, P, device=dev), torch.randn(TP, P, device=dev), torch.randn(P, device=dev)
    timeit(name, lambda: ops.gemm(a, wt, cview(out), bias=bias, r0=cview(r)), gflop=2.0 * TP * P * K / 1e9)

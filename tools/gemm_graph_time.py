"""GPU-side time per launch of small / medium GEMM shapes, host overhead excluded: 20 launches captured in one CUDA graph,
replayed 5 times (CUDA events). Shows the fixed cost per launch (time at 1 wave) against the per-wave cost.
usage: python tools/gemm_graph_time.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rosettafold_pytorch_b200 import ops
from rosettafold_pytorch_b200.ops import cview

dev = torch.device("cuda:0")
REP = 20


def run(name, T, N, K, odt, res=False, bias=True):
    x = (torch.randn(T, K, device=dev) * 0.1).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.1).bfloat16()
    out = torch.empty(T, N, dtype=odt, device=dev)
    b = torch.randn(N, device=dev) if bias else None
    r = torch.randn(T, N, device=dev) if res else None

    def launch():
        ops.gemm(x, w, cview(out), bias=b, r0=None if r is None else cview(r))
    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REP):
            launch()
    g.replay()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(e) * 1e3 / (5 * REP)
    print(f"{name:34s} T={T:6d} N={N:5d} K={K:5d}  {us:7.1f} us  {2.0 * T * N * K / us / 1e6:7.1f} TFLOP/s", flush=True)


bf, f32 = torch.bfloat16, torch.float32
for T in (2048, 8192, 18944, 37888, 65536):
    run("plain 384->384 bf16", T, 384, 384, bf)
for T in (2048, 18944, 65536):
    run("to_out 384->384 +res f32", T, 384, 384, f32, res=True)
for T in (2048, 18944, 65536):
    run("tied [q|pk] 384->768 bf16", T, 768, 384, bf)
for T in (2048, 18944, 65536):
    run("FF1 384->1536 bf16", T, 1536, 384, bf)
for T in (2048, 65536):
    run("FF2 1536->384 +res f32", T, 384, 1536, f32, res=True)

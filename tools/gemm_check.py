"""bf16-output GEMMs (ragged M, K tails, the big projection shapes) vs a float32 matmul of the same bf16 operands on
sampled rows, plus timing. usage: python tools/gemm_check.py  (A/B: RFK_GEMM_BN, RFK_GEMM_NO_TMA_EPILOGUE)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rosettafold_pytorch_b200 import ops
from rosettafold_pytorch_b200.ops import cview

dev = torch.device("cuda:0")
for T, N, K, act in [(4096, 256, 64, 0), (4224, 256, 104, 1), (8192 + 128, 512, 288, 0), (65536, 768, 384, 0),
                     (65536, 2304, 384, 0), (262144, 1536, 288, 0), (262144, 1152, 288, 1)]:
    torch.manual_seed(T + N)
    x = (torch.randn(T, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.5).bfloat16()
    b = torch.randn(N, device=dev)
    out = torch.empty(T, N, dtype=torch.bfloat16, device=dev)
    ops.gemm(x, w, cview(out), bias=b, act=act)
    torch.cuda.synchronize()
    rows = torch.randint(0, T, (2048,), device=dev)
    rows[:256] = torch.arange(T - 256, T, device=dev)   # the ragged last tile
    ref = x[rows].float() @ w.float().t() + b
    if act:
        ref = ref.relu()
    err = float((out[rows].float() - ref).norm() / ref.norm())
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        ops.gemm(x, w, cview(out), bias=b, act=act)
    e.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(e) * 100
    print(f"T={T:7d} N={N:5d} K={K:4d} rel-l2 {err:.2e}  {us:8.1f} us  {2.0 * T * N * K / us / 1e6:7.1f} TFLOP/s", flush=True)

"""Calibration only (never on the product path): cuBLAS (torch.nn.functional.linear, bf16 in / bf16 out) on the big
projection shapes of the trunk, for comparison with tools/gemm_check.py. usage: python tools/gemm_cublas.py"""
import torch
import torch.nn.functional as F

dev = torch.device("cuda:0")
for T, N, K in [(65536, 768, 384), (65536, 2304, 384), (65536, 1536, 384), (262144, 1536, 288), (262144, 1152, 288),
                (262144, 288, 1152), (262144, 288, 512)]:
    x = (torch.randn(T, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.5).bfloat16()
    b = torch.randn(N, device=dev).bfloat16()
    for _ in range(3):
        F.linear(x, w, b)
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        F.linear(x, w, b)
    e.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(e) * 100
    print(f"cuBLAS T={T:7d} N={N:5d} K={K:4d} bf16 out {us:8.1f} us  {2.0 * T * N * K / us / 1e6:7.1f} TFLOP/s", flush=True)

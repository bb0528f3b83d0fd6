"""One GEMM shape, a few launches (ncu target). usage: one_gemm.py T N K out_dtype(bf16|f32) res(0|1)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rosettafold_pytorch_b200 import ops
from rosettafold_pytorch_b200.ops import cview
T, N, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
odt = torch.bfloat16 if sys.argv[4] == "bf16" else torch.float32
res = len(sys.argv) > 5 and sys.argv[5] == "1"
dev = torch.device("cuda:0")
x = (torch.randn(T, K, device=dev) * 0.1).bfloat16(); w = (torch.randn(N, K, device=dev) * 0.1).bfloat16()
out = torch.empty(T, N, dtype=odt, device=dev); b = torch.randn(N, device=dev)
r = torch.randn(T, N, device=dev) if res else None
for _ in range(4):
    ops.gemm(x, w, cview(out), bias=b, r0=None if r is None else cview(r))
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))

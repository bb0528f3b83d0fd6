"""One FAVOR launch (ncu target): usage one_favor.py G T H kind"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rosettafold_pytorch_b200 import ops
G, T, H, kind = (int(a) for a in sys.argv[1:5])
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
proj = torch.randn(266, 64, generator=g).to(dev)
inner = H * 64
buf = (torch.randn(1, G, T, 3 * inner, generator=g) * 0.7).to(torch.bfloat16).to(dev)
out = torch.zeros(1, G, T, inner, dtype=torch.bfloat16, device=dev)
for _ in range(3):
    ops.favor_attention(buf[..., :inner], buf[..., inner:2 * inner], buf[..., 2 * inner:], out, proj, kind=kind, heads=H)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))

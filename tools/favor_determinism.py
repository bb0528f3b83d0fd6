"""Run the FAVOR kernel on interleaved shapes and report history dependence of the outputs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rosettafold_pytorch_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
proj = torch.randn(266, 64, generator=g).to(dev)
shapes = [(54, 7, 12, 0, True), (54, 18, 8, 1, True), (54, 18, 8, 1, False), (18, 7, 12, 0, True), (18, 18, 8, 1, True), (18, 18, 8, 1, False)]
bufs = {}
for s in shapes:
    G, T, H, kind, strided = s
    inner = H * 64
    if strided:
        bufs[s] = (torch.randn(1, T, G, 3 * inner, generator=g) * 0.7).to(torch.bfloat16).to(dev).permute(0, 2, 1, 3)
    else:
        bufs[s] = (torch.randn(1, G, T, 3 * inner, generator=g) * 0.7).to(torch.bfloat16).to(dev)
def run(s):
    G, T, H, kind, strided = s
    inner = H * 64
    buf = bufs[s]
    out = torch.full((1, T, G, inner) if strided else (1, G, T, inner), float("nan"), dtype=torch.bfloat16, device=dev)
    if strided: out = out.permute(0, 2, 1, 3)
    ops.favor_attention(buf[..., :inner], buf[..., inner:2 * inner], buf[..., 2 * inner:], out, proj, kind=kind, heads=H)
    return out.float()
first = {s: run(s) for s in shapes}
for rnd in range(3):
    for s in shapes[::-1] if rnd % 2 else shapes:
        o = run(s)
        torch.cuda.synchronize()
        d = (o - first[s]).abs()
        print(rnd, s, "nan:", int(torch.isnan(o).sum()), "max diff", float(d.max()), "n diff", int((d > 0).sum()))

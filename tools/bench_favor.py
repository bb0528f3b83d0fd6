"""FAVOR kernel accuracy + timing at the metric shapes. Run on the GPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rosettafold_pytorch_b200 import ops
from oracle.ops_ref import RefBackend
REF = RefBackend()
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
proj = torch.randn(266, 64, generator=g).to(dev)

def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())

def case(name, G1, G0, T, H, kind, token_dim_strided, check=True, reps=5, scale=0.7):
    inner = H * 64
    if token_dim_strided:
        buf = (torch.randn(G1, T, G0, 3 * inner, generator=g) * scale).to(torch.bfloat16).to(dev)
        view = buf.permute(0, 2, 1, 3)
        out = torch.zeros(G1, T, G0, inner, dtype=torch.bfloat16, device=dev).permute(0, 2, 1, 3)
    else:
        buf = (torch.randn(G1, G0, T, 3 * inner, generator=g) * scale).to(torch.bfloat16).to(dev)
        view = buf
        out = torch.zeros(G1, G0, T, inner, dtype=torch.bfloat16, device=dev)
    q, k, v = view[..., :inner], view[..., inner:2 * inner], view[..., 2 * inner:]
    ops.favor_attention(q, k, v, out, proj, kind=kind, heads=H)
    torch.cuda.synchronize()
    msg = ""
    if check:
        ref = torch.empty(out.shape, dtype=torch.float32, device=dev)
        REF.favor_attention(q, k, v, ref, proj, kind, H)
        msg = f"rel-l2 {rel(out, ref):.3e}"
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        ops.favor_attention(q, k, v, out, proj, kind=kind, heads=H)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    fl = 8.0 * G1 * G0 * T * H * 64 * 266
    print(f"{name:40s} {ms*1e3:9.1f} us {fl/ms/1e9:8.1f} TFLOP/s  {msg}")

case("small relu T=40", 1, 3, 40, 2, 1, True)
case("small softmax T=40", 1, 3, 40, 2, 0, True)
case("relu T=200 strided", 1, 2, 200, 2, 1, True)
case("softmax T=200 strided", 1, 2, 200, 2, 0, True)
case("softmax T=128 G=64 H=12", 1, 64, 128, 12, 0, True)
case("relu T=512 G=32 H=8", 1, 32, 512, 8, 1, False)
case("MSA col: softmax G=512 T=128 H=12", 1, 512, 128, 12, 0, True, check=False)
case("pair row: relu G=512 T=512 H=8 strided", 1, 512, 512, 8, 1, True, check=False)
case("pair col: relu G=512 T=512 H=8", 1, 512, 512, 8, 1, False, check=False)

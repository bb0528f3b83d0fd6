"""One launch of the MSA-column FAVOR shape (softmax kernel, G=512, T=128, H=12) for the -DRFK_COL_TIMELINE build of
librfk (tools/build_variant.sh tl -DRFK_COL_TIMELINE; RFK_LIB_PATH=.../librfk_tl.so): the kernel prints its stamps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rosettafold_pytorch_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
proj = torch.randn(266, 64, generator=g).to(dev)
G0, T, H = 512, 128, 12
inner = H * 64
buf = (torch.randn(1, T, G0, 3 * inner, generator=g) * 0.7).to(torch.bfloat16).to(dev)
view = buf.permute(0, 2, 1, 3)
out = torch.zeros(1, T, G0, inner, dtype=torch.bfloat16, device=dev).permute(0, 2, 1, 3)
q, k, v = view[..., :inner], view[..., inner:2 * inner], view[..., 2 * inner:]
ops.favor_attention(q, k, v, out, proj, kind=0, heads=H)
torch.cuda.synchronize()

"""One launch of the pair-axis FAVOR shape (ReLU kernel, G=512, T=512, H=8) for the -DRFK_TM_TIMELINE build of librfk
(tools/build_variant.sh tmtl -DRFK_TM_TIMELINE; RFK_LIB_PATH=.../librfk_tmtl.so): the kernel prints its stamps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rosettafold_pytorch_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
proj = torch.randn(266, 64, generator=g).to(dev)
G0, T, H = 512, 512, 8
inner = H * 64
buf = (torch.randn(1, G0, T, 3 * inner, generator=g) * 0.7).to(torch.bfloat16).to(dev)
out = torch.zeros(1, G0, T, inner, dtype=torch.bfloat16, device=dev)
q, k, v = buf[..., :inner], buf[..., inner:2 * inner], buf[..., 2 * inner:]
ops.favor_attention(q, k, v, out, proj, kind=1, heads=H)
torch.cuda.synchronize()

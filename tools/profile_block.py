"""One TwoTrackBlock at the metric config, twice (ncu launch-list target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rosettafold_pytorch_b200 as rf
dev = torch.device("cuda:0")
N, L = int(os.environ.get("N", 128)), int(os.environ.get("L", 512))
torch.manual_seed(0)
blk = rf.TwoTrackBlock(384, 288, n_encoder_layers=4).eval().to(dev)
msa = torch.randn(1, N, L, 384, device=dev); pair = torch.randn(1, L, L, 288, device=dev)
for i in range(2):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); m, p = blk(msa, pair); b.record(); torch.cuda.synchronize()
    print("block ms", a.elapsed_time(b))

"""Host-side cost of enqueueing one TwoTrackBlock: at a tiny size the GPU work is negligible, so the
event time per block is the Python/ctypes/tensor-map launch overhead (developer tool)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rosettafold_pytorch_b200 as rf
dev = torch.device("cuda:0")
torch.manual_seed(0)
blk = rf.TwoTrackBlock(384, 288, n_encoder_layers=4).eval().to(dev)
for (N, L) in [(16, 64), (128, 512)]:
    msa = torch.randn(1, N, L, 384, device=dev); pair = torch.randn(1, L, L, 288, device=dev)
    for _ in range(3): blk(msa, pair)
    torch.cuda.synchronize()
    n0 = rf._lib.launch_count()
    t0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): m, p = blk(msa, pair)
    b.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"N={N} L={L}: enqueue {1e3*(t1-t0)/5:.2f} ms/block (host), device {a.elapsed_time(b)/5:.2f} ms/block, "
          f"wall {1e3*(t2-t0)/5:.2f} ms/block, {(rf._lib.launch_count()-n0)//5} librfk launches/block")
# the same block through a CUDA graph
g = rf.GraphedModule(blk)
for (N, L) in [(16, 64), (64, 256), (128, 512)]:
    msa = torch.randn(1, N, L, 384, device=dev); pair = torch.randn(1, L, L, 288, device=dev)
    for _ in range(3): blk(msa, pair)
    g(msa, pair); torch.cuda.synchronize()
    res = {}
    for name, fn in (("eager", blk), ("graph", g)):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): fn(msa, pair)
        b.record(); torch.cuda.synchronize()
        res[name] = a.elapsed_time(b) / 5
    print(f"N={N} L={L}: eager {res['eager']:.2f} ms/block, graph {res['graph']:.2f} ms/block")

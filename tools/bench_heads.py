"""PredictionHead (4 ResNets x 4 dilated residual blocks, d_pair 288) at L = 512: time per forward (CUDA events, eager and
as one CUDA graph) and parity against the golden fixture. Run on the GPU box: python tools/bench_heads.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rosettafold_pytorch_b200 as rf
from oracle.weights import synth_state_dict
from tests.helpers import build_heads, load_golden, rel_l2

dev = torch.device("cuda:0")
for name in ("small", "default"):
    gold = load_golden("prediction_head")[name]
    head, _, pair = build_heads(gold["config"], dev)
    for mode in ("fp32", "bf16"):
        rf.set_mode(mode)
        out = head(pair.to(dev))
        torch.cuda.synchronize()
        print(name, mode, {k: f"{rel_l2(out[k], gold[k]):.2e}" for k in out})
rf.set_mode("bf16")
L, C = 512, 288
head = rf.PredictionHead(C, 4, 0.1).eval()
head.load_state_dict(synth_state_dict(head.state_dict(), seed=5))
head = head.to(dev)
pair = torch.randn(1, L, L, C, device=dev)
n0 = rf._lib.launch_count()
head(pair)
torch.cuda.synchronize()
launches = rf._lib.launch_count() - n0
head(pair)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    head(pair)
b.record()
torch.cuda.synchronize()
eager = a.elapsed_time(b) / 3
g = rf.GraphedModule(head)
g(pair); g(pair)
a.record()
for _ in range(3):
    g(pair)
b.record()
torch.cuda.synchronize()
graph = a.elapsed_time(b) / 3
flops = 4 * (2.0 * L * L * C * C * (1 + 9 * 8)) + 2.0 * L * L * C * (C + 37 * 3 + 19)
print(f"PredictionHead L={L}: {launches} librfk launches, eager {eager:.2f} ms, graph {graph:.2f} ms, "
      f"{flops / graph / 1e9:.0f} TFLOP/s on the contractions")

"""Find the first librfk call whose outputs differ between two identical TwoTrackBlock forwards
(developer tool): wraps every backend method, snapshots all tensor arguments after each call."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rosettafold_pytorch_b200 as rf
from rosettafold_pytorch_b200 import ops
from tests.helpers import build_block
dev = torch.device("cuda:0")
cfg = dict(d_msa=96, d_pair=72, n_layers=1, B=3, N=7, L=18, seed=5)
blk, _, msa, pair = build_block(cfg, dev)
be = ops.backend()
log = []
def wrap(name, fn):
    def f(*a, **k):
        pre = [x.clone() if isinstance(x, torch.Tensor) else None for x in a]
        r = fn(*a, **k)
        torch.cuda.synchronize()
        log.append((name, pre, [x.clone() if isinstance(x, torch.Tensor) else None for x in a]))
        return r
    return f
for n in dir(be):
    if not n.startswith("_") and callable(getattr(be, n)) and n not in ("lib",):
        setattr(be, n, wrap(n, getattr(be, n)))
def run():
    log.clear()
    blk(msa, pair)
    return list(log)
run()
a = run()
blk(msa[1:2].contiguous(), pair[1:2].contiguous())
b = run()
print(len(a), len(b))
for i, ((n1, pre1, post1), (n2, pre2, post2)) in enumerate(zip(a, b)):
    din = [float((x.float() - y.float()).abs().nan_to_num(1e9).max()) for x, y in zip(pre1, pre2) if x is not None and x.shape == y.shape]
    dout = [float((x.float() - y.float()).abs().nan_to_num(1e9).max()) for x, y in zip(post1, post2) if x is not None and x.shape == y.shape]
    if max(dout + [0]) > 0:
        print(i, n1, "inputs differ:", din, "outputs differ:", dout, [tuple(x.shape) for x in post1 if x is not None])
        if sum(1 for _ in range(1)) and i > 0 and max(din + [0]) == 0:
            print("   ^ first divergence with identical inputs")
        if i > 400: break

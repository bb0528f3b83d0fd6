"""Developer tool: per-stage parity of a TwoTrackBlock against the CPU oracle at given shapes.
usage: python tools/parity_stages.py [mode] B,N,L,layers [B,N,L,layers ...]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rosettafold_pytorch_b200 as rf  # noqa: E402
from oracle import trunk_ref  # noqa: E402
from tests.helpers import STAGES, build_block, rel_l2, run_stages  # noqa: E402

args = sys.argv[1:]
mode = "bf16"
if args and args[0] in ("bf16", "fp32"):
    mode = args.pop(0)
rf.set_mode(mode)
dev = torch.device("cuda:0")
for spec in args:
    B, N, L, nl = (int(x) for x in spec.split(","))
    cfg = dict(d_msa=384, d_pair=288, n_layers=nl, B=B, N=N, L=L, seed=41)
    blk, sd, msa, pair = build_block(cfg, dev)
    t0 = time.time()
    gold = {}
    with torch.no_grad():
        trunk_ref.two_track_block(msa.cpu(), pair.cpu(), sd, nl, stages=gold)
    forced = run_stages(blk, msa, pair, teacher=gold)
    chain = run_stages(blk, msa, pair)
    torch.cuda.synchronize()
    print(mode, spec, "oracle %.1fs" % (time.time() - t0))
    print("  forced", {k: "%.2e" % rel_l2(forced[k], gold[k]) for k in STAGES})
    print("  chain ", {k: "%.2e" % rel_l2(chain[k], gold[k]) for k in STAGES}, flush=True)

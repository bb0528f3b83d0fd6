#!/usr/bin/env python
"""Long-protein path (BASELINE.json config 4): one TwoTrackBlock (or, with --blocks K, a trunk of K blocks) at
(1, N, L), sharded over the ranks as described in rosettafold_pytorch_b200/sharded.py, checked against the
single-GPU run on rank 0 and timed on the device (CUDA events, max over ranks).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29533 tools/bench_long.py --L 1024 --N 256 [--blocks 2]
  RFK_SHARD_TIMING=1 ... prints the device time of every stage of every block (adds a synchronize per block).
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import rosettafold_pytorch_b200 as rf

ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, default=1024)
ap.add_argument("--N", type=int, default=256)
ap.add_argument("--layers", type=int, default=4)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--blocks", type=int, default=1, help="> 1: a trunk of that many blocks; the pair map stays "
                "row-sharded between blocks (ShardedTrunkBlocks), reported per block")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)  # identical weights (and FAVOR projections) on every rank
if args.blocks > 1:
    blk = rf.TrunkBlocks(384, 288, n_blocks=args.blocks, n_encoder_layers=args.layers).eval().to(dev)
else:
    blk = rf.TwoTrackBlock(384, 288, n_encoder_layers=args.layers).eval().to(dev)
g = torch.Generator().manual_seed(5)
msa = torch.randn((1, args.N, args.L, 384), generator=g).to(dev)
pair = torch.randn((1, args.L, args.L, 288), generator=g).to(dev)
sblk = rf.ShardedTrunkBlocks(blk) if args.blocks > 1 else rf.ShardedTwoTrackBlock(blk)

def timed(fn):
    for _ in range(2): fn()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps): out = fn()
    b.record(); torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / args.steps], device=dev)
    if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms) / args.blocks, out

ms_sh, (m_sh, p_sh) = timed(lambda: sblk(msa, pair))
res = {"config": f"{args.blocks} x TwoTrackBlock (1,{args.N},{args.L}), {args.layers} encoder layers, row-/sequence-sharded x{world} (tied row layers by sequence, Performer column layers by residue, pair stages by row)",
       "n_gpus": world, "ms_per_block_sharded": ms_sh}
if rank == 0:
    ms_1, (m_1, p_1) = timed(lambda: blk(msa, pair)) if world == 1 else (None, blk(msa, pair))
    torch.cuda.synchronize()
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    res.update(rel_l2_msa_vs_single_gpu=rel(m_sh, m_1), rel_l2_pair_vs_single_gpu=rel(p_sh, p_1))
if world > 1:
    dist.barrier()
    # single-GPU timing of the same block on every rank (max), for the speed-up
    ms_1, _ = timed(lambda: blk(msa, pair))
    res["ms_per_block_single_gpu"] = ms_1
    res["speedup"] = ms_1 / ms_sh
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()

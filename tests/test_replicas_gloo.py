"""N>1 path on CPU: world_size-2 gloo processes run the trunk as batch replicas (oracle-emulated
ops) and must reproduce the single-process result; covers even and ragged batch splits."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, batch, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import rosettafold_pytorch_b200 as rf
    from oracle.ops_ref import RefBackend
    from rosettafold_pytorch_b200 import ops
    from rosettafold_pytorch_b200.replicas import run_replicated
    from tests.helpers import build_block

    ops._set_backend_for_tests(RefBackend())
    rf.set_mode("fp32")
    cfg = dict(d_msa=48, d_pair=40, n_layers=1, B=batch, N=4, L=10, seed=9)
    blk, _, msa, pair = build_block(cfg)
    m, p = run_replicated(blk, msa, pair)
    if rank == 0:
        m1, p1 = blk(msa, pair)
        torch.save(dict(err_m=float((m - m1).abs().max()), err_p=float((p - p1).abs().max()),
                        shape=tuple(m.shape)), os.path.join(out_dir, f"res_{batch}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [2, 3])
def test_two_rank_replicas_match_single_process(tmp_path, batch):
    port = 29500 + (os.getpid() % 2000) + batch
    # spawned children re-import this module by name: make the repo root importable for them
    os.environ["PYTHONPATH"] = ROOT + os.pathsep + os.environ.get("PYTHONPATH", "")
    mp.spawn(_worker, args=(2, port, batch, str(tmp_path)), nprocs=2, join=True)
    res = torch.load(os.path.join(tmp_path, f"res_{batch}.pt"))
    assert res["shape"][0] == batch
    assert res["err_m"] < 1e-5 and res["err_p"] < 1e-5, res


def test_shard_bounds_cover_batch():
    from rosettafold_pytorch_b200.replicas import shard_bounds

    for B in (1, 2, 3, 7, 64):
        for w in (1, 2, 4, 8):
            cuts = [shard_bounds(B, r, w) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == B
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in cuts) - min(h - l for l, h in cuts) <= 1

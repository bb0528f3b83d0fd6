"""Long-protein path on CPU: world_size-2 gloo processes run the sharded stages (pair axial attention with the
all-to-all transpose, sequence-sharded tied row layers, whole block; oracle-emulated ops) and must reproduce
the single-process result."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import rosettafold_pytorch_b200 as rf
    from oracle.ops_ref import RefBackend
    from rosettafold_pytorch_b200 import ops
    from rosettafold_pytorch_b200.sharded import (ShardedPairAxialAttention, ShardedTwoTrackBlock,
                                                    all_to_all_cols_to_rows, all_to_all_rows_to_cols, row_shard)
    from tests.helpers import build_block

    ops._set_backend_for_tests(RefBackend())
    rf.set_mode("fp32")
    cfg = dict(d_msa=48, d_pair=40, n_layers=2, B=1, N=6, L=12, seed=13)
    blk, _, msa, pair = build_block(cfg)
    L = cfg["L"]
    lo, hi = row_shard(L, rank, world)
    # 1. the two transposes are inverse permutations of the pair tensor
    x = pair[0, lo:hi].contiguous()                                   # [Li, L, D]
    cols = all_to_all_rows_to_cols(x, None)                           # [L, Lj, D]
    err_t = float((cols - pair[0][:, lo:hi]).abs().max())
    back = all_to_all_cols_to_rows(cols, None)                        # [P, Li, Lj, D]
    err_b = float((back.permute(1, 0, 2, 3).reshape(hi - lo, L, -1) - x).abs().max())
    # 2. sharded axial stage == rows of the unsharded one
    ax = ShardedPairAxialAttention(blk.pair_update_with_axial_attention)
    rows = ax(pair[:, lo:hi].contiguous())
    ref = blk.pair_update_with_axial_attention(pair)
    err_ax = float((rows - ref[:, lo:hi]).abs().max())
    # 3. MSA self-attention: tied row layers sequence-sharded (broadcast query row, merged position-wise softmax,
    #    all-reduced logits), all-to-all, Performer column layers residue-sharded
    sblk = ShardedTwoTrackBlock(blk)
    n_lo, n_hi = row_shard(cfg["N"], rank, world)
    ms, atts = sblk._msa_self_attention(msa[:, n_lo:n_hi].contiguous(), rank, world)   # -> residue shard
    mr, attr = blk.msa_update_using_self_att(msa)
    err_sa = max(float((ms - mr[:, :, lo:hi]).abs().max()), float((atts - attr).abs().max()))
    # 4. whole block, outputs replicated on every rank
    m, p = sblk(msa, pair)
    m1, p1 = blk(msa, pair)
    # 5. two blocks: the pair map stays row-sharded between them (MsaUpdateWithPair's attention maps come from the
    #    row shard + its all-to-all transpose)
    from rosettafold_pytorch_b200.sharded import ShardedTrunkBlocks
    from oracle.weights import synth_state_dict
    trunk = rf.TrunkBlocks(cfg["d_msa"], cfg["d_pair"], n_blocks=2, n_encoder_layers=1).eval()
    trunk.load_state_dict(synth_state_dict(trunk.state_dict(), seed=17), strict=True)
    mt, pt = ShardedTrunkBlocks(trunk)(msa, pair)
    mt1, pt1 = trunk(msa, pair)
    err_trunk = max(float((mt - mt1).abs().max()), float((pt - pt1).abs().max()))
    torch.save(dict(err_t=err_t, err_b=err_b, err_ax=err_ax, err_sa=err_sa, err_m=float((m - m1).abs().max()),
                    err_p=float((p - p1).abs().max()), err_trunk=err_trunk), os.path.join(out_dir, f"res_{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])  # with three ranks the middle one has a halo neighbour on both sides
def test_sharded_stages_match_single_process(tmp_path, world):
    port = 31500 + (os.getpid() % 2000) + world
    os.environ["PYTHONPATH"] = ROOT + os.pathsep + os.environ.get("PYTHONPATH", "")
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for rank in range(world):
        res = torch.load(os.path.join(tmp_path, f"res_{rank}.pt"))
        assert res["err_t"] == 0.0 and res["err_b"] == 0.0, res
        assert res["err_ax"] < 1e-4 and res["err_sa"] < 1e-4 and res["err_m"] < 1e-4 and res["err_p"] < 1e-4, res
        assert res["err_trunk"] < 1e-4, res


def test_row_shard_rejects_ragged():
    from rosettafold_pytorch_b200.sharded import row_shard

    assert row_shard(12, 1, 3) == (4, 8)
    with pytest.raises(ValueError):
        row_shard(10, 0, 4)

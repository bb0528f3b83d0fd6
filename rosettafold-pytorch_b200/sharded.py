"""One long protein over several GPUs (BASELINE.json config 4).

Between the stages of a block - and between the blocks of a trunk (`ShardedTrunkBlocks`) - no tensor of the
size of the MSA or of the pair map is ever replicated: the MSA lives as sequence shards `[1, N/P, L, D]` or
residue shards `[1, N, L/P, D]`, the pair map as row shards `[1, L/P, L, D]` (rank r owns rows
[r L/P, (r+1) L/P)). Per block (reference :962-968):

  A `MsaUpdateUsingSelfAttention`
    * tied row layers on the sequence shard. Projections, the q scaling, A.V, to_out and the FeedForward are
      per sequence; the layer couples sequences in three places only (`SequenceShard`): the query row of the
      position-wise weights is sequence 0 of the whole MSA (broadcast from rank 0: L x d_msa operand-dtype
      elements), their softmax runs over all sequences (merged from per-rank (max, sum) statistics: an
      all-gather of 2 x L x 12 floats) and the logits sum over all sequences (all-reduce of 12 x L x L floats);
    * one all-to-all turns sequence shards into residue shards;
    * Performer column layers (tokens along n, one group per residue) on the residue shard: local.
  B `PairUpdateWithMsa` on the pair rows: the MSA enters through its 32-channel projection only, computed on the
    residue shard and all-gathered (N L 32 floats); outer-product sum, 716-wide Linear and the convolution block
    run on the rank's rows. Each 3x3 convolution needs one neighbour row on either side (an all-gather of every
    rank's first and last row: 2 L d_pair elements per rank), the InstanceNorm statistics are all-reduced
    (2 x d_pair doubles).
  C `PairUpdateWithAxialAttention` on the pair rows: column attention (tokens along j), every LayerNorm and the
    FeedForward are local; row attention (tokens along i) needs whole columns: the normalised operand is
    transposed between ranks with ONE all-to-all (rank r receives columns [r L/P, (r+1) L/P) of every row),
    attention + output projection run on the column shard, and a second all-to-all brings the update back to the
    row shard where it is added to the fp32 residual stream. Only operand-dtype tensors (bf16 in the tensor-core
    mode) cross NVLink: 2 x L^2 D / P elements per rank and layer.
  D `MsaUpdateWithPair`: one all-to-all turns the residue shards back into sequence shards (every op is
    independent per sequence). The attention maps need the symmetrised pair map (:555-556), the one place where
    pair rows meet their columns: one all-to-all delivers each rank the columns of its rows, it computes the
    logits of its rows (`rfk_pair2att_logits_rows`) and the ranks all-gather the 16 x L x L logits.

`ShardedTwoTrackBlock.forward` / `ShardedTrunkBlocks.forward` take and return replicated tensors (they shard at
entry and all-gather once at exit); `forward_rows` is the shard-to-shard form. N and L must be divisible by the
number of ranks. One process per GPU; `torch.distributed` (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist
import torch.nn as nn

from . import modules as M


_STAGE_TIMING = bool(os.environ.get("RFK_SHARD_TIMING"))  # developer aid: per-stage device times of every block

# ---------------------------------------------------------------------------------------------------------------------
# Collectives go through `_collective`, which is what lets `SegmentedGraph` run a sharded module as CUDA graphs of the
# compute segments with the NCCL calls issued eagerly in between (NCCL itself is never captured).
# ---------------------------------------------------------------------------------------------------------------------
_recorder = None


def _collective(fn):
    """Run one collective (`fn()` = a torch.distributed call on tensors that already exist). Under a SegmentedGraph
    recording the current capture is cut here and `fn` is kept for the replays."""
    if _recorder is None:
        fn()
    else:
        _recorder.cut(fn)


class SegmentedGraph(nn.Module):
    """CUDA-graph execution of a module whose forward mixes librfk kernels with torch.distributed collectives.

    A sharded block issues ~200 kernel launches and ~40 collectives; at 8 ranks each GPU has ~10 ms of work per
    block while the host needs ~14 ms to enqueue it through Python: the sharded path is host-bound exactly where it is
    supposed to scale. The first call with a given input shape runs the module once under a recorder: everything
    between two collectives is captured into its own CUDA graph (all graphs share one memory pool, so the tensors that
    cross a cut keep their addresses), every collective is executed eagerly and its closure kept. Later calls copy
    the inputs into the static input buffers and replay graph, collective, graph, ... in the recorded order: one
    graph launch per segment instead of one Python round trip per kernel. Every rank records and replays the same
    sequence (the control flow of the sharded modules depends on shapes only). The outputs are views of graph-owned
    buffers, overwritten by the next call."""

    def __init__(self, module: nn.Module, warmup: int = 2):
        super().__init__()
        self.module = module
        self.warmup = warmup
        self._plans = {}

    def _record(self, inputs):
        global _recorder
        static_in = [t.clone() for t in inputs]
        with torch.no_grad():
            for _ in range(self.warmup):  # communicators, kernel attributes, weight packing: all outside the capture
                self.module(*static_in)
        torch.cuda.synchronize()
        plan = self
        graphs, colls = [], []

        class Rec:
            def __init__(self):
                self.pool = torch.cuda.graph_pool_handle()
                self.cur = None

            def begin(self):
                self.cur = torch.cuda.CUDAGraph()
                self.cur.capture_begin(pool=self.pool)

            def cut(self, fn):
                self.cur.capture_end()
                graphs.append(self.cur)
                fn()  # (moves whatever the buffers hold: the captured kernels have not run; only the order matters)
                colls.append(fn)
                self.begin()

            def end(self):
                self.cur.capture_end()
                graphs.append(self.cur)

        side = torch.cuda.Stream(device=static_in[0].device)
        side.wait_stream(torch.cuda.current_stream())
        rec = Rec()
        with torch.cuda.stream(side), torch.no_grad():
            _recorder = rec
            try:
                rec.begin()
                out = self.module(*static_in)
                rec.end()
            finally:
                _recorder = None
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        del plan
        return static_in, out, graphs, colls

    @torch.no_grad()
    def forward(self, *inputs):
        key = tuple((tuple(t.shape), t.dtype, t.device) for t in inputs) + (M.get_mode(), M._BOUNDED_DT)
        plan = self._plans.get(key)
        if plan is None:
            plan = self._plans[key] = self._record(inputs)
        static_in, out, graphs, colls = plan
        for s_, t in zip(static_in, inputs):
            s_.copy_(t, non_blocking=True)
        for i, g in enumerate(graphs):
            g.replay()
            if i < len(colls):
                colls[i]()
        return out

    def segments(self):
        """(number of graphs, number of collectives) of the recorded plans (diagnostics)."""
        return [(len(p[2]), len(p[3])) for p in self._plans.values()]


def _as_like(t):
    return t if t.dtype == torch.float32 else t.float()


def row_shard(L: int, rank: int, world: int):
    if L % world:
        raise ValueError(f"pair row-sharding needs L ({L}) divisible by the number of ranks ({world})")
    n = L // world
    return rank * n, (rank + 1) * n


def all_to_all_rows_to_cols(x_rows: torch.Tensor, group=None) -> torch.Tensor:
    """[Li, L, D] (my rows, all columns) -> [L, Lj, D] (all rows, my columns). One all-to-all."""
    world = dist.get_world_size(group)
    Li, L, D = x_rows.shape
    Lj = L // world
    # chunk p = my rows x the columns of rank p
    send = x_rows.view(Li, world, Lj, D).permute(1, 0, 2, 3).contiguous()
    recv = torch.empty_like(send)  # chunk p = rows of rank p x my columns: [P, Li, Lj, D] == [L, Lj, D]
    _collective(lambda: dist.all_to_all_single(recv, send, group=group))
    return recv.view(world * Li, Lj, D)


def all_to_all_cols_to_rows(x_cols: torch.Tensor, group=None) -> torch.Tensor:
    """[L, Lj, D] (all rows, my columns) -> [P, Li, Lj, D]: chunk p holds my rows x the columns of
    rank p (the caller scatters it into its [Li, L, D] layout). One all-to-all, no packing."""
    world = dist.get_world_size(group)
    L, Lj, D = x_cols.shape
    Li = L // world
    send = x_cols.contiguous().view(world, Li, Lj, D)  # chunk p = rows of rank p
    recv = torch.empty_like(send)
    _collective(lambda: dist.all_to_all_single(recv, send, group=group))
    return recv


class SequenceShard:
    """The collectives of a tied row layer whose sequences are split over the ranks of `group`
    (hooks called by `SoftTiedAttentionOverResidues._attend`)."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.root = dist.get_global_rank(group, 0) if group is not None else 0

    def first_sequence(self, xn4: torch.Tensor) -> torch.Tensor:
        """[B, N/P, L, D] -> [B, 1, L, D]: sequence 0 of the whole MSA (rank 0's first row)."""
        row = xn4[:, :1].clone()  # (a size-1 slice is "contiguous": .contiguous() would alias xn4 and the broadcast overwrite it)
        _collective(lambda: dist.broadcast(row, src=self.root, group=self.group))
        return row

    def softmax_correction(self, stats: torch.Tensor) -> torch.Tensor:
        """stats [B, L, H, 2] = (max, sum of exp) of this rank's position-wise logits -> the factor [B, L, H]
        that turns weights normalised over this rank's sequences into weights normalised over all of them."""
        allst = torch.empty((self.world,) + tuple(stats.shape), dtype=stats.dtype, device=stats.device)
        dst, src = allst.view(-1), stats.reshape(-1)
        _collective(lambda: dist.all_gather_into_tensor(dst, src, group=self.group))
        gmax = allst[..., 0].max(dim=0).values
        gsum = (allst[..., 1] * torch.exp(allst[..., 0] - gmax)).sum(dim=0)
        return stats[..., 1] * torch.exp(stats[..., 0] - gmax) / gsum

    def allreduce(self, t: torch.Tensor) -> None:
        _collective(lambda: dist.all_reduce(t, group=self.group))

    def reduce_scatter_rows(self, t: torch.Tensor) -> torch.Tensor:
        """[R, C] contiguous partial sums on every rank -> [R / P, C]: rows [rank R / P, (rank + 1) R / P) of the sum."""
        out = torch.empty((t.shape[0] // self.world, t.shape[1]), dtype=t.dtype, device=t.device)
        src = t.reshape(-1)
        _collective(lambda: dist.reduce_scatter_tensor(out.view(-1), src, group=self.group))
        return out

    def all_gather_rows(self, mine: torch.Tensor, full: torch.Tensor) -> None:
        """[R / P, C] (this rank's rows) -> full [R, C] contiguous, rows in rank order."""
        dst, src = full.view(-1), mine.reshape(-1)
        _collective(lambda: dist.all_gather_into_tensor(dst, src, group=self.group))


def all_to_all_seqs_to_residues(x_seqs: torch.Tensor, group=None) -> torch.Tensor:
    """[1, N/P, L, D] (my sequences, all residues) -> [1, N, L/P, D] (all sequences, my residues)."""
    world = dist.get_world_size(group)
    _, Nl, L, D = x_seqs.shape
    Ll = L // world
    send = x_seqs.view(Nl, world, Ll, D).permute(1, 0, 2, 3).contiguous()  # chunk p = my sequences x residues of p
    recv = torch.empty_like(send)                                           # chunk q = sequences of q x my residues
    _collective(lambda: dist.all_to_all_single(recv, send, group=group))
    return recv.view(1, world * Nl, Ll, D)


def all_to_all_residues_to_seqs(x_res: torch.Tensor, group=None) -> torch.Tensor:
    """[1, N, L/P, D] (all sequences, my residues) -> [1, N/P, L, D] (my sequences, all residues)."""
    world = dist.get_world_size(group)
    _, N, Ll, D = x_res.shape
    Nl = N // world
    send = x_res.contiguous().view(world, Nl, Ll, D)       # chunk p = sequences of rank p x my residues
    recv = torch.empty_like(send)                          # chunk q = my sequences x residues of rank q
    _collective(lambda: dist.all_to_all_single(recv, send, group=group))
    return recv.permute(1, 0, 2, 3).reshape(1, Nl, world * Ll, D)


class ShardedPairAxialAttention(nn.Module):
    """`PairUpdateWithAxialAttention` on a row shard of the pair tensor (B = 1)."""

    def __init__(self, module: M.PairUpdateWithAxialAttention, group=None):
        super().__init__()
        self.module = module
        self.group = group

    @torch.no_grad()
    def forward(self, x_rows: torch.Tensor) -> torch.Tensor:
        """x_rows: float32 [1, Li, L, D], rows [rank Li, (rank+1) Li) of the pair tensor."""
        if x_rows.dim() != 4 or x_rows.shape[0] != 1:
            raise ValueError("ShardedPairAxialAttention: expects one protein, x_rows [1, L/P, L, D]")
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if world == 1:
            return self.module(x_rows)
        x = x_rows.float().clone(memory_format=torch.contiguous_format)  # updated in place below
        _, Li, L, D = x.shape
        if Li * world != L:
            raise ValueError(f"row shard of {Li} rows x {world} ranks does not cover L = {L}")
        adt = M._adt()
        for layer in self.module.layers:
            # ---- row attention (tokens along i): transpose the normalised operand between ranks ----
            x2 = x.view(Li * L, D)
            xn = M._ln_into(x2, layer.layer[0].fn[0], M._empty(x2.shape, adt, x2))
            xt = all_to_all_rows_to_cols(xn.view(Li, L, D), self.group)            # [L, Lj, D]
            upd = layer.row_attn._run(xt.view(1, L, L // world, D), 1, None, out_dtype=adt)  # [L*Lj, D]
            back = all_to_all_cols_to_rows(upd.view(L, L // world, D), self.group)  # [P, Li, Lj, D]
            # residual add in the row layout: x[i, p Lj + jj, :] += back[p, i, jj, :]
            x.view(Li, world, L // world, D).add_(back.permute(1, 0, 2, 3))
            # ---- column attention and feed-forward are local to the row shard ----
            x = M._performer_block(layer.layer[1].fn[0], layer.col_attn, x, token_dim=2)
            x = M._ff_block(layer.layer[2].fn[0], layer.ff, x.view(-1, D), adt).view(1, Li, L, D)
        return x


class ShardedTwoTrackBlock(nn.Module):
    """A `TwoTrackBlock` for one long protein on `world` GPUs: replicated MSA track and
    PairUpdateWithMsa, row-sharded pair axial attention, all-gather of the pair rows."""

    def __init__(self, block: M.TwoTrackBlock, group=None):
        super().__init__()
        self.block = block
        self.axial = ShardedPairAxialAttention(block.pair_update_with_axial_attention, group)
        self.group = group

    def _msa_self_attention(self, msa_seq, rank, world):
        """MsaUpdateUsingSelfAttention (:399-409): msa_seq [1, N/P, L, D] = this rank's sequences -> the updated
        MSA as this rank's residues [1, N, L/P, D], and the (replicated) symmetrised tied attention map. Tied row
        layers on the sequence shard, one all-to-all, Performer column layers on the residue shard."""
        mod = self.block.msa_update_using_self_att
        xq = M._as_f32(msa_seq).contiguous()
        row_shard(xq.shape[2], rank, world)                   # L must divide too
        att = None
        n = len(mod.residue_wise_encoder_layers)
        shard = SequenceShard(self.group)
        for i, layer in enumerate(mod.residue_wise_encoder_layers):
            xq, a = layer._run(xq, want_att=(i == n - 1), shard=shard)
            att = a if a is not None else att
        xs = all_to_all_seqs_to_residues(xq, self.group)      # [1, N, L/P, D]
        for layer in mod.sequence_wise_encoder_layers:
            xs, _ = layer._run(xs, token_dim=1)
        return xs, att

    def _pair_update_with_msa(self, msa_res, rows, att, rank, world):
        """PairUpdateWithMsa (:465-498) for this rank's rows [1, L/P, L, d_pair] of the pair map. msa_res
        [1, N, L/P, D]: the MSA enters through its 32-channel projection only (per token), which is computed on
        the residue shard and all-gathered (N L 32 floats instead of N L 384)."""
        lo, hi = row_shard(rows.shape[2], rank, world)
        group = self.group
        mod = self.block.pair_update_with_msa
        part = mod._project(msa_res)                                             # [1, N, L/P, Q]
        allm = torch.empty((world,) + tuple(part.shape), dtype=part.dtype, device=part.device)
        dst, src = allm.view(-1), part.reshape(-1)
        _collective(lambda: dist.all_gather_into_tensor(dst, src, group=group))
        mraw = allm.permute(1, 2, 0, 3, 4).reshape(part.shape[0], part.shape[1], -1, part.shape[3])  # [1, N, L, Q]

        def halo(x):  # [1, Li, L, C] -> [1, Li + 2, L, C]
            edges = torch.stack([x[:, 0], x[:, -1]], 0).contiguous()            # my first / last row
            allv = torch.empty((world,) + tuple(edges.shape), dtype=x.dtype, device=x.device)
            dst, src = allv.view(-1), edges.view(-1)
            _collective(lambda: dist.all_gather_into_tensor(dst, src, group=group))
            zero = torch.zeros_like(x[:, :1])
            top = allv[rank - 1, 1].unsqueeze(1) if rank > 0 else zero          # last row of the rank above
            bottom = allv[rank + 1, 0].unsqueeze(1) if rank + 1 < world else zero
            return torch.cat([top, x, bottom], 1)

        def allreduce(st):
            _collective(lambda: dist.all_reduce(st, group=group))

        return mod._forward_rows(None, rows, att[:, lo:hi], lo, hi, halo, allreduce, mraw=mraw)

    def _msa_update_with_pair(self, msa_res, rows, rank, world):
        """MsaUpdateWithPair (:607-610): msa_res [1, N, L/P, D] -> this rank's sequences [1, N/P, L, D] of the
        updated MSA (every op is independent per sequence; one all-to-all turns residue shards into sequence
        shards). The attention maps come from the row-sharded pair map: the symmetrisation (:555-556) needs each
        row's columns, which one all-to-all of the row shards delivers (L^2 d_pair / P floats per rank instead
        of an all-gather of the whole map); every rank then computes the logits of its rows and the ranks
        all-gather the 16 x L x L logits."""
        group = self.group
        msa_seq = all_to_all_residues_to_seqs(msa_res, group)
        rows = rows.contiguous()
        cols_t = all_to_all_rows_to_cols(rows[0], group).unsqueeze(0)          # [1, L, L/P, D]

        def gather_rows(part):  # [1, C, Li, L] -> [1, C, L, L]
            allp = torch.empty((world,) + tuple(part.shape), dtype=part.dtype, device=part.device)
            dst, src = allp.view(-1), part.reshape(-1)
            _collective(lambda: dist.all_gather_into_tensor(dst, src, group=group))
            return allp.permute(1, 2, 0, 3, 4).reshape(part.shape[0], part.shape[1], -1, part.shape[3])

        mod = self.block.msa_update_with_pair
        return mod._run(msa_seq, lambda chunk: M._pair2att_rows(chunk, rows, cols_t, gather_rows))

    @torch.no_grad()
    def forward_rows(self, msa_seq: torch.Tensor, rows: torch.Tensor):
        """msa_seq [1,N/P,L,d_msa] = this rank's sequences, rows [1,L/P,L,d_pair] = this rank's rows of the pair
        map -> the same shards of the block's outputs. Neither tensor is ever gathered: consecutive blocks hand
        the shards to each other (`ShardedTrunkBlocks`)."""
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if msa_seq.shape[0] != 1:
            raise ValueError("ShardedTwoTrackBlock: one protein per call (batches run as replicas)")
        marks = [] if _STAGE_TIMING and msa_seq.is_cuda else None

        def mark(name):
            if marks is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append((name, ev))

        mark("start")
        msa_res, att = self._msa_self_attention(msa_seq, rank, world)
        mark("msa_self_attention")
        rows = self._pair_update_with_msa(msa_res, _as_like(rows), att, rank, world)
        mark("pair_update_with_msa")
        rows = self.axial(rows)
        mark("pair_axial_attention")
        msa_seq = self._msa_update_with_pair(msa_res, rows, rank, world)
        mark("msa_update_with_pair")
        if marks is not None:
            torch.cuda.synchronize()
            print(f"[rank {rank}] " + ", ".join(f"{n} {a.elapsed_time(b):.2f} ms"
                                                for (_, a), (n, b) in zip(marks[:-1], marks[1:])), flush=True)
        return msa_seq, rows

    @torch.no_grad()
    def forward(self, msa: torch.Tensor, pair: torch.Tensor):
        """msa [1,N,L,d_msa], pair [1,L,L,d_pair] replicated on every rank -> same, replicated."""
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if world == 1:
            return self.block(msa, pair)
        msa_seq, rows = shard_inputs(msa, pair, self.group)
        msa_seq, rows = self.forward_rows(msa_seq, rows)
        return gather_shards(msa_seq, self.group), gather_shards(rows, self.group)


def shard_inputs(msa: torch.Tensor, pair: torch.Tensor, group=None):
    """Replicated msa [1,N,L,D], pair [1,L,L,P] -> (this rank's sequences, this rank's pair rows)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_lo, n_hi = row_shard(msa.shape[1], rank, world)
    lo, hi = row_shard(pair.shape[1], rank, world)
    return msa[:, n_lo:n_hi].contiguous(), pair[:, lo:hi].contiguous()


def gather_shards(x: torch.Tensor, group=None) -> torch.Tensor:
    """[1, X/P, ...] shards of axis 1 on every rank -> the whole [1, X, ...] tensor on every rank."""
    world = dist.get_world_size(group)
    full = torch.empty((x.shape[0], x.shape[1] * world) + tuple(x.shape[2:]), dtype=x.dtype, device=x.device)
    dst, src = full.view(-1), x.reshape(-1)
    _collective(lambda: dist.all_gather_into_tensor(dst, src, group=group))
    return full


class ShardedTrunkBlocks(nn.Module):
    def __init__(self, trunk: M.TrunkBlocks, group=None):
        super().__init__()
        self.blocks = nn.ModuleList([ShardedTwoTrackBlock(b, group) for b in trunk.blocks])

    @torch.no_grad()
    def forward(self, msa, pair):
        """Replicated in, replicated out; between the blocks the MSA stays sequence-sharded and the pair map
        row-sharded."""
        if not self.blocks:
            return msa, pair
        group = self.blocks[0].group
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if world == 1:
            for blk in self.blocks:
                msa, pair = blk(msa, pair)
            return msa, pair
        msa_seq, rows = shard_inputs(msa, pair, group)
        for blk in self.blocks:
            msa_seq, rows = blk.forward_rows(msa_seq, rows)
        return gather_shards(msa_seq, group), gather_shards(rows, group)

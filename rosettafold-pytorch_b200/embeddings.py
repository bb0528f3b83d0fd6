"""Drop-in mirrors of the reference's embeddings (rosettafold_pytorch/rosettafold_pytorch.py:57-181): the callers on
the input side of the trunk.

Same class names, constructor / forward signatures and state_dict keys as the reference. What changes:
  * device placement: the sinusoidal tables are (non-persistent) buffers, so `.to(device)` moves them and state_dict
    keys stay the reference's; nothing is gathered by Python loops over the batch (:73, :98) or built on the CPU
    (:115-116) — the reference as published cannot run these modules on a GPU;
  * MsaEmbedding / PairEmbedding forward = ONE fused gather kernel each (rfk_msa_embed / rfk_pair_embed). The Linear of
    PairEmbedding (:173) acts on a concatenation of two gathered embeddings and one scalar feature: it is applied to the
    21-row embedding table at weight-packing time instead, so the (B, L, L, 289) concatenation of :171 never exists.
    With a template, its LayerNorm + the template columns of the Linear run on the trunk's LayerNorm / GEMM kernels.
Inference only (dropout is the identity), float32 outputs in the trunk's layouts.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import modules as M
from . import ops


def _sinusoid_table(dim: int, max_len: int) -> torch.Tensor:
    """reference :63-68 / :86-91"""
    pe = torch.zeros(max_len, dim)
    denom = torch.exp(math.log(10000.0) * torch.arange(0, dim, 2) / dim)
    pos = torch.arange(0, max_len).view(-1, 1)
    pe[:, 0::2] = torch.sin(pos / denom)
    pe[:, 1::2] = torch.cos(pos / denom)
    return pe


class SinusoidalPositionalEncoding(nn.Module):
    """reference :57-77. API-parity helper (the fused path adds the table inside rfk_msa_embed)."""

    def __init__(self, dim, max_len, p_dropout=0.1):
        super().__init__()
        self.dim, self.max_len = dim, max_len
        self.register_buffer("pos_enc", _sinusoid_table(dim, max_len), persistent=False)
        self.dropout = nn.Dropout(p_dropout)

    @torch.no_grad()
    def forward(self, x, aa_idx):
        return x + self.pos_enc[aa_idx.to(self.pos_enc.device)].unsqueeze(1)


class SinusoidalPositionalEncoding2D(nn.Module):
    """reference :79-103. API-parity helper (the fused path adds the table inside rfk_pair_embed)."""

    def __init__(self, dim, max_len, p_dropout=0.1):
        super().__init__()
        self.max_len = max_len
        self.register_buffer("pos_enc", _sinusoid_table(dim // 2, max_len), persistent=False)
        self.dropout = nn.Dropout(p_dropout)

    @torch.no_grad()
    def forward(self, x, aa_idx):
        L = aa_idx.size(1)
        pe = self.pos_enc[aa_idx.to(self.pos_enc.device)]
        return x + torch.cat([pe[:, :, None, :].expand(-1, -1, L, -1), pe[:, None, :, :].expand(-1, L, -1, -1)], dim=-1)


class MsaEmbedding(nn.Module):
    """reference :106-120"""

    def __init__(self, d_input=21, d_msa=384, max_len=260, p_pe_drop=0.1):
        super().__init__()
        self.to_embedding = nn.Embedding(d_input, d_msa)
        self.pos_enc = SinusoidalPositionalEncoding(d_msa, max_len, p_pe_drop)
        self.query_enc = nn.Embedding(2, d_msa)  # 0: query, 1: targets

    @torch.no_grad()
    def forward(self, x, aa_idx):
        """x: (B, N, L) int64 tokens, aa_idx: (B, L) int64 residue indices -> (B, N, L, d_msa) float32."""
        dev = self.to_embedding.weight.device
        x, aa_idx = x.to(dev).contiguous(), aa_idx.to(dev).contiguous()
        B, N, L = x.shape
        out = torch.empty((B, N, L, self.to_embedding.embedding_dim), dtype=torch.float32, device=dev)
        return ops.msa_embed(x, aa_idx, M._f(self.to_embedding.weight), self.pos_enc.pos_enc, M._f(self.query_enc.weight), out)


class PairEmbedding(nn.Module):
    """reference :123-181"""

    def __init__(self, d_input=21, d_pair=288, max_len=260, p_pe_drop=0.1, use_template=False, d_template=64):
        super().__init__()
        self.half_d_pair = d_pair // 2
        self.embed_seq = nn.Embedding(d_input, self.half_d_pair)
        self.pos_enc = SinusoidalPositionalEncoding2D(d_pair, max_len, p_pe_drop)
        self.use_template = use_template
        if self.use_template:
            self.ln_template = nn.LayerNorm(d_template)
            self.proj = nn.Linear(d_pair + d_template + 1, d_pair)
        else:
            self.proj = nn.Linear(d_pair + 1, d_pair)

    def _pack(self):
        def build():
            h = self.half_d_pair
            W = self.proj.weight.detach().double()
            emb = self.embed_seq.weight.detach().double()
            d = dict(
                # the Linear of :173 applied to the two gathered halves of the concatenation (:171): per-vocabulary tables
                left=(emb @ W[:, :h].T).float().contiguous(),         # "b l d -> b k l d": residue j
                right=(emb @ W[:, h:2 * h].T).float().contiguous(),   # "b l d -> b l k d": residue i
                sep=W[:, 2 * h].float().contiguous(), bias=M._f(self.proj.bias))
            if self.use_template:
                d["Wt"] = M._w(self.proj.weight[:, 2 * h + 1:], M._bdt())  # the operand is a LayerNorm output
            return d
        return M._packed(self, build)

    @torch.no_grad()
    def forward(self, seq, aa_idx, template=None):
        if not self.use_template and template is not None:
            raise ValueError(f"[{self.__class__.__name__}]: template is not None but use_template is False")
        dev = self.proj.weight.device
        seq, aa_idx = seq.to(dev).contiguous(), aa_idx.to(dev).contiguous()
        pk = self._pack()
        B, L = seq.shape
        D = self.proj.out_features
        out = torch.empty((B, L, L, D), dtype=torch.float32, device=dev)
        ops.pair_embed(seq, aa_idx, pk["left"], pk["right"], pk["sep"], pk["bias"], self.pos_enc.pos_enc, out)
        if self.use_template:
            # + W_t LayerNorm(template) (:160-167): LayerNorm kernel + GEMM with the gathered part as the residual
            t = M._as_f32(template).to(dev).contiguous()
            dt = t.shape[-1]
            tn = M._ln_into(t.view(-1, dt), self.ln_template, torch.empty((B * L * L, M._up8(dt)), dtype=M._bdt(), device=dev)[:, :dt])
            res = out
            out = torch.empty_like(res)
            ops.gemm(tn, pk["Wt"], ops.cview(out.view(-1, D)), r0=ops.cview(res.view(-1, D)))
        return out

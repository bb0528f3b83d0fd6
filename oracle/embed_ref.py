"""ORACLE / TEST INFRASTRUCTURE — not part of the product.

CPU restatement (plain PyTorch, functional) of the reference's embeddings, the callers on the input side of the trunk
(SURVEY.md section 8(f) rank 3): /root/reference/rosettafold_pytorch/rosettafold_pytorch.py:57-181. Weights come in as a
flat dict with the reference's state_dict key names. Pinned by tests/golden/embeddings.pt (outputs of the UNMODIFIED
reference, oracle/make_golden.py --embeddings-only) and by a live comparison in tests/test_oracle.py.
"""
import math

import torch


def sinusoid_table(dim, max_len):
    """:63-68 / :86-91 — pos_enc[p, 2k] = sin(p / 10000^(2k/dim)), pos_enc[p, 2k+1] = cos(same)."""
    pe = torch.zeros(max_len, dim)
    denom = torch.exp(math.log(10000.0) * torch.arange(0, dim, 2) / dim)
    pos = torch.arange(0, max_len).view(-1, 1)
    pe[:, 0::2] = torch.sin(pos / denom)
    pe[:, 1::2] = torch.cos(pos / denom)
    return pe


def msa_embedding(tokens, aa_idx, sd, max_len, prefix=""):
    """MsaEmbedding.forward :114-120 (dropout inert in eval): tokens (B,N,L) int64, aa_idx (B,L) int64 -> (B,N,L,d_msa)."""
    emb, qenc = sd[prefix + "to_embedding.weight"], sd[prefix + "query_enc.weight"]
    pe = sinusoid_table(emb.shape[1], max_len)[aa_idx]  # (B, L, D)  (:73)
    query_idx = torch.ones(tokens.shape[-2], 1, dtype=torch.long)  # (:115-116)
    query_idx[0] = 0
    return (emb[tokens] + pe[:, None]) + qenc[query_idx]


def pair_embedding(seq, aa_idx, sd, max_len, template=None, prefix=""):
    """PairEmbedding.forward :147-175: seq, aa_idx (B,L) int64 -> (B,L,L,d_pair)."""
    emb = sd[prefix + "embed_seq.weight"]
    W, b = sd[prefix + "proj.weight"], sd[prefix + "proj.bias"]
    B, L = seq.shape
    d_pair = W.shape[0]
    e = emb[seq]  # (B, L, d_pair/2)
    left = e[:, None, :, :].expand(B, L, L, -1)   # "b l d -> b k l d": residue j
    right = e[:, :, None, :].expand(B, L, L, -1)  # "b l d -> b l k d": residue i
    sep = torch.log((aa_idx[:, :, None] - aa_idx[:, None, :]).abs() + 1)[..., None]  # :177-181
    parts = [left, right, sep]
    if template is not None:
        parts.append(torch.nn.functional.layer_norm(template, (template.shape[-1],), sd[prefix + "ln_template.weight"],
                                                    sd[prefix + "ln_template.bias"], 1e-5))
    x = torch.nn.functional.linear(torch.cat(parts, dim=-1), W, b)
    pe_half = sinusoid_table(d_pair // 2, max_len)[aa_idx]  # (B, L, d_pair/2)  (:98)
    pe = torch.cat([pe_half[:, :, None, :].expand(B, L, L, -1), pe_half[:, None, :, :].expand(B, L, L, -1)], dim=-1)
    return x + pe

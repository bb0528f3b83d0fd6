"""ORACLE / TEST INFRASTRUCTURE — not part of the product.  *** PARITY UNPINNED ***

Restatement of `performer_pytorch.SelfAttention` as the reference calls it
(rosettafold_pytorch.py:10, :313-318, :505-518). The dependency is un-vendored and un-pinned
(`setup.py:24` lists "performer-pytorch" with no version); it is absent from /root/reference,
from this image and there is no network, so this file restates the published algorithm of
lucidrains/performer-pytorch 1.1.4 (what an unpinned install resolves to): FAVOR+ random
features (Choromanski et al. 2021) with
  * dim_head = 64 regardless of dim/heads, nb_features = int(64*ln 64) = 266,
  * q/k/v projections without bias, output projection with bias,
  * Gaussian-orthogonal projection matrix drawn once at construction (buffer),
  * softmax kernel (eps 1e-4, per-query / global-key max stabilisers) or, with
    generalized_attention=True, the ReLU kernel (eps 1e-3),
  * non-causal linear attention  out = (q' (k'^T v)) / (q' . sum_n k').
No golden vector for this arithmetic exists in the reference (tests pin shapes only,
tests/test_module.py:266-278, :390-402): parity for this component is anchored on the
reference's call sites alone.
"""
import math

import torch
from torch import nn

DIM_HEAD = 64


def nb_features_default(dim_head=DIM_HEAD):
    return int(dim_head * math.log(dim_head))


def orthogonal_block(cols, generator=None):
    """One square block with orthonormal rows: Q^T of a QR of a Gaussian matrix."""
    g = torch.randn((cols, cols), generator=generator)
    q, _ = torch.linalg.qr(g, mode="reduced")
    return q.t()


def gaussian_orthogonal_random_matrix(nb_rows, nb_cols, generator=None):
    """Stacked orthogonal blocks, rows rescaled by norms of Gaussian vectors (scaling = 0)."""
    blocks = [orthogonal_block(nb_cols, generator) for _ in range(nb_rows // nb_cols)]
    rest = nb_rows - (nb_rows // nb_cols) * nb_cols
    if rest:
        blocks.append(orthogonal_block(nb_cols, generator)[:rest])
    mat = torch.cat(blocks)
    norms = torch.randn((nb_rows, nb_cols), generator=generator).norm(dim=1)
    return norms[:, None] * mat


def softmax_features(x, proj, is_query, eps=1e-4):
    """x: (..., n, d) -> positive random features (..., n, m) approximating exp(q.k)."""
    d = x.shape[-1]
    c = d ** -0.25
    u = torch.einsum("...nd,md->...nm", c * x, proj.to(x.dtype))
    half_sq = (x * x).sum(-1, keepdim=True) * (0.5 * c * c)
    stab = u.amax(dim=-1, keepdim=True) if is_query else u.amax(dim=(-1, -2), keepdim=True)
    return (proj.shape[0] ** -0.5) * (torch.exp(u - half_sq - stab) + eps)


def relu_features(x, proj, eps=1e-3):
    c = x.shape[-1] ** -0.25
    return torch.relu(torch.einsum("...nd,md->...nm", c * x, proj.to(x.dtype))) + eps


def linear_attention(qf, kf, v):
    ksum = kf.sum(dim=-2)
    denom = torch.einsum("...nm,...m->...n", qf, ksum)
    ctx = torch.einsum("...nm,...ne->...me", kf, v)
    return torch.einsum("...nm,...me->...ne", qf, ctx) / denom[..., None]


def favor_attention(q, k, v, proj, generalized):
    """q, k, v: (b, h, n, 64)."""
    if generalized:
        return linear_attention(relu_features(q, proj), relu_features(k, proj), v)
    return linear_attention(softmax_features(q, proj, True), softmax_features(k, proj, False), v)


class FastAttention(nn.Module):
    def __init__(self, dim_heads, nb_features=None, generalized_attention=False):
        super().__init__()
        self.nb_features = nb_features if nb_features is not None else nb_features_default(dim_heads)
        self.generalized_attention = generalized_attention
        self.register_buffer("projection_matrix",
                             gaussian_orthogonal_random_matrix(self.nb_features, dim_heads))

    def forward(self, q, k, v):
        return favor_attention(q, k, v, self.projection_matrix, self.generalized_attention)


class SelfAttention(nn.Module):
    """Constructor surface the reference uses: (dim, heads, dropout, generalized_attention)."""

    def __init__(self, dim, causal=False, heads=8, dim_head=DIM_HEAD, nb_features=None,
                 generalized_attention=False, dropout=0.0, qkv_bias=False, attn_out_bias=True, **unused):
        super().__init__()
        if causal or unused:
            raise NotImplementedError(f"oracle restates the non-causal default path only: {unused}")
        inner = dim_head * heads
        self.heads = heads
        self.fast_attention = FastAttention(dim_head, nb_features, generalized_attention)
        self.to_q = nn.Linear(dim, inner, bias=qkv_bias)
        self.to_k = nn.Linear(dim, inner, bias=qkv_bias)
        self.to_v = nn.Linear(dim, inner, bias=qkv_bias)
        self.to_out = nn.Linear(inner, dim, bias=attn_out_bias)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x):
        b, n, _ = x.shape
        h = self.heads

        def heads_first(t):
            return t.reshape(b, n, h, -1).transpose(1, 2)

        out = self.fast_attention(heads_first(self.to_q(x)), heads_first(self.to_k(x)),
                                  heads_first(self.to_v(x)))
        return self.dropout(self.to_out(out.transpose(1, 2).reshape(b, n, -1)))

"""ORACLE / TEST INFRASTRUCTURE — not part of the product.

CPU restatement (plain PyTorch, functional, fp32 or fp64) of the reference's trunk path:
`TwoTrackBlock.forward` and everything below it. Weights come in as a flat dict with the
reference's state_dict key names (plus `msa_update_with_pair.encoder_layers.{i}.*` for the
layers the reference keeps in a plain list). Every function cites the lines of
/root/reference/rosettafold_pytorch/rosettafold_pytorch.py it follows.

Pinned by: tests/golden/*.pt (outputs of the UNMODIFIED reference run in the build container by
oracle/make_golden.py) and, when /root/reference is present, a live comparison
(tests/test_oracle.py). The Performer arithmetic inside is `oracle/performer_ref.py`
— PARITY UNPINNED for that dependency (see its header).

Used by: tests/ (checker), __graft_entry__.smoke() (checker) and bench.py's cpu_baseline /
`--impl reference` legs (timed CPU baseline). Never imported by the product package.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .performer_ref import favor_attention


class W:
    """Prefix view over a flat weight dict."""

    def __init__(self, sd, prefix=""):
        self.sd, self.prefix = sd, prefix

    def sub(self, name):
        return W(self.sd, f"{self.prefix}{name}.")

    def __getitem__(self, name):
        return self.sd[self.prefix + name]

    def has(self, name):
        return (self.prefix + name) in self.sd


def linear(x, w: W, name):
    return F.linear(x, w[f"{name}.weight"], w[f"{name}.bias"] if w.has(f"{name}.bias") else None)


def layer_norm(x, w: W, name, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), w[f"{name}.weight"], w[f"{name}.bias"], eps)


def feed_forward(x, w: W):
    """:270-281 (dropout inert in eval)"""
    return linear(torch.relu(linear(x, w, "net.0")), w, "net.3")


def poswise_weight(x, w: W, n_heads):
    """:205-217 — x (b n l d) -> softmax over n of <to_q(x[:,0])*scale, to_k(x)> per (b,l,h)."""
    b, n, l, d = x.shape
    dh = d // n_heads
    q = linear(x[:, 0], w, "to_q.0").reshape(b, l, n_heads, dh) * dh ** -0.5
    k = linear(x, w, "to_k.0").reshape(b, n, l, n_heads, dh)
    logits = torch.einsum("blhd,bnlhd->blhn", q, k)
    return torch.softmax(logits, dim=-1).permute(0, 3, 2, 1).unsqueeze(-1)  # b n h l 1


def tied_attention(x, w: W, n_heads):
    """:241-267 — returns (to_out(att . v), symmetrised att as (b i j h))."""
    b, n, l, d = x.shape
    dh = d // n_heads

    def heads(t):
        return t.reshape(b, n, l, n_heads, dh).permute(0, 1, 3, 2, 4)  # b n h l d

    q, k, v = heads(linear(x, w, "to_q")), heads(linear(x, w, "to_k")), heads(linear(x, w, "to_v"))
    q = q * poswise_weight(x, w.sub("poswise_weight"), n_heads) * dh ** -0.5
    att = torch.softmax(torch.einsum("bnhid,bnhjd->bhij", q, k), dim=-1)
    out = torch.einsum("bhij,bnhjd->bnhid", att, v).permute(0, 1, 3, 2, 4).reshape(b, n, l, d)
    sym = (0.5 * (att + att.transpose(-1, -2))).permute(0, 2, 3, 1)
    return linear(out, w, "to_out"), sym


def performer_attention(x, w: W, heads, generalized):
    """performer_pytorch.SelfAttention.forward on x (batch, tokens, dim) — :313-318, :505-518."""
    bsz, n, _ = x.shape

    def split(t):
        return t.reshape(bsz, n, heads, 64).transpose(1, 2)

    q, k, v = split(linear(x, w, "to_q")), split(linear(x, w, "to_k")), split(linear(x, w, "to_v"))
    out = favor_attention(q, k, v, w["fast_attention.projection_matrix"].to(x.dtype), generalized)
    return linear(out.transpose(1, 2).reshape(bsz, n, heads * 64), w, "to_out")


def encoder_layer_tied(x, w: W, n_heads):
    """:334-354 with tied=True."""
    a, att = tied_attention(layer_norm(x, w, "ln"), w.sub("attn"), n_heads)
    x = x + a
    return x + feed_forward(layer_norm(x, w, "ff.fn.0"), w.sub("ff.fn.1")), att


def encoder_layer_performer(x, w: W, n_heads):
    """:334-354 with performer=True: x (b n l d) is flattened to ((b n) l d) (:338)."""
    b, n, l, d = x.shape
    a = performer_attention(layer_norm(x, w, "ln").reshape(b * n, l, d), w.sub("attn"), n_heads, False)
    x = x + a.reshape(b, n, l, d)
    return x + feed_forward(layer_norm(x, w, "ff.fn.0"), w.sub("ff.fn.1"))


def msa_update_using_self_att(x, w: W, n_layers, n_heads=12):
    """:399-409"""
    att = None
    for i in range(n_layers):
        x, att = encoder_layer_tied(x, w.sub(f"residue_wise_encoder_layers.{i}"), n_heads)
    x = x.transpose(1, 2)  # b l n d (:403)
    for i in range(n_layers):
        x = encoder_layer_performer(x, w.sub(f"sequence_wise_encoder_layers.{i}"), n_heads)
    return x.transpose(1, 2), att


def outer_product_mean(x, y, w: W):
    """:419-427 — a SUM over n, then LayerNorm(u*v) and Linear."""
    o = torch.einsum("bniu,bnjv->bijuv", x, y).flatten(-2)
    return linear(layer_norm(o, w, "to_out.0"), w, "to_out.1")


def pair_update_with_msa(msa, pair, att, w: W):
    """:465-498"""
    L = msa.shape[2]
    m = layer_norm(linear(layer_norm(msa, w, "proj_msa.0"), w, "proj_msa.1"), w, "proj_msa.2")
    wt = poswise_weight(m, w.sub("poswise_weight"), 1)[:, :, 0]  # b n l 1
    coevol = layer_norm(outer_product_mean(m, m * wt, w.sub("outer_product_mean")), w, "ln_coevol_feat")
    msa_1d = torch.cat([m.sum(1), m[:, 0]], dim=-1)  # b l 2q
    feat = torch.cat([
        coevol,
        msa_1d[:, :, None, :].expand(-1, -1, L, -1),   # feature of row index i
        msa_1d[:, None, :, :].expand(-1, L, -1, -1),   # feature of column index j
        layer_norm(pair, w, "ln_pair"),
        att,
    ], dim=-1)
    h = linear(feat, w, "resnet.0")
    y = h.permute(0, 3, 1, 2)
    y = F.conv2d(y, w["resnet.1.fn.1.weight"], padding=1)
    y = F.elu(F.instance_norm(y, weight=w["resnet.1.fn.2.weight"], bias=w["resnet.1.fn.2.bias"], eps=1e-6))
    y = F.conv2d(y, w["resnet.1.fn.5.weight"], padding=1)
    y = F.instance_norm(y, weight=w["resnet.1.fn.6.weight"], bias=w["resnet.1.fn.6.bias"], eps=1e-6)
    return F.elu(y.permute(0, 2, 3, 1) + h)


def pair_axial_layer(x, w: W, n_heads=8):
    """:521-525 — row attention (over axis 1), column attention (over axis 2), feed-forward."""
    b, n, l, d = x.shape
    xn = layer_norm(x, w, "layer.0.fn.0").transpose(1, 2).reshape(b * l, n, d)  # (b l) n d (:51)
    x = x + performer_attention(xn, w.sub("row_attn"), n_heads, True).reshape(b, l, n, d).transpose(1, 2)
    xn = layer_norm(x, w, "layer.1.fn.0").reshape(b * n, l, d)  # (b n) l d (:38)
    x = x + performer_attention(xn, w.sub("col_attn"), n_heads, True).reshape(b, n, l, d)
    return x + feed_forward(layer_norm(x, w, "layer.2.fn.0"), w.sub("ff"))


def pair_update_with_axial_attention(x, w: W, n_layers):
    """:544-547"""
    for i in range(n_layers):
        x = pair_axial_layer(x, w.sub(f"layers.{i}"))
    return x


def msa_update_with_pair_layer(msa, pair, w: W, n_heads=4):
    """:588-595"""
    b, n, l, d = msa.shape
    sym = 0.5 * (pair + pair.transpose(1, 2))
    att = torch.softmax(linear(layer_norm(sym, w, "pair2att.1"), w, "pair2att.2").permute(0, 3, 1, 2), dim=-1)
    v = linear(layer_norm(msa, w, "msa2value.0"), w, "msa2value.1").reshape(b, n, l, n_heads, d // n_heads)
    upd = torch.einsum("bhij,bnjhd->bnihd", att, v).reshape(b, n, l, d)
    y = msa + upd
    return y + feed_forward(layer_norm(y, w, "ff.fn.0"), w.sub("ff.fn.1"))


def msa_update_with_pair(msa, pair, w: W, n_layers):
    """:607-610 — every layer sees the same pair."""
    for i in range(n_layers):
        msa = msa_update_with_pair_layer(msa, pair, w.sub(f"encoder_layers.{i}"))
    return msa


def msa_update_with_pair_and_coord(xyz, state, msa, w: W, distance_bins=(8, 12, 16, 20), ca_idx=1):
    """:889-920 — distance-masked attention from the state track applied to the normalised MSA."""
    h = len(distance_bins)
    b, n, l, d = msa.shape
    state = layer_norm(state, w, "ln_state")
    msa = layer_norm(msa, w, "ln_msa")
    q = linear(state, w, "to_q").reshape(b, l, h, -1).permute(0, 2, 1, 3)
    k = linear(state, w, "to_k").reshape(b, l, h, -1).permute(0, 2, 1, 3)
    v = linear(msa, w, "to_v").reshape(b, n, l, h, d // h)
    scale = (state.shape[-1] // h) ** -0.5
    pdist = torch.cdist(xyz[:, :, ca_idx], xyz[:, :, ca_idx])
    mask = torch.stack([(pdist < t).to(msa.dtype) for t in distance_bins], dim=1)
    logits = torch.einsum("bhid,bhjd->bhij", q * scale, k) + (1.0 - mask) * -1e9
    att = logits.softmax(dim=-1)
    out = torch.einsum("bhij,bnjhd->bnihd", att, v).reshape(b, n, l, d)
    msa = msa + layer_norm(out, w, "ln_out")
    return msa + feed_forward(layer_norm(msa, w, "to_out.fn.0"), w.sub("to_out.fn.1"))


def two_track_block(msa, pair, sd, n_layers, prefix="", stages=None):
    """:962-968. `stages`, if a dict, receives the intermediate tensors."""
    w = W(sd, prefix)
    msa, att = msa_update_using_self_att(msa, w.sub("msa_update_using_self_att"), n_layers)
    p1 = pair_update_with_msa(msa, pair, att, w.sub("pair_update_with_msa"))
    p2 = pair_update_with_axial_attention(p1, w.sub("pair_update_with_axial_attention"), n_layers)
    m2 = msa_update_with_pair(msa, p2, w.sub("msa_update_with_pair"), n_layers)
    if stages is not None:
        stages.update(msa_a=msa, att=att, pair_b=p1, pair_c=p2, msa_d=m2)
    return m2, p2

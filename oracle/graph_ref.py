"""ORACLE / TEST INFRASTRUCTURE — not part of the product.

CPU restatement of the reference's dense graph transformer: `GraphTransformer.forward` (rosettafold_pytorch.py:632-664) and
`GraphTransformerBlock.forward` (:676-677), functional over a state_dict with the reference's keys, eval mode. Pinned by
tests/golden/graph_transformer.pt (outputs of the UNMODIFIED reference, oracle/make_golden.py --graph-only) in
tests/test_oracle.py.
"""
import torch
import torch.nn.functional as F


def graph_transformer(node, edge, mask, w, n_heads):
    """:632-664. node [B,L,Dn], edge [B,L,L,De], mask [B,L,L] (1 = edge exists) or None -> [B, L, H*d]."""
    B, L, _ = node.shape
    H = n_heads

    def lin(name, x, bias=True):
        return F.linear(x, w[f"{name}.weight"], w[f"{name}.bias"] if bias else None)

    def heads(t):                                                   # b l (h d) -> b h l d   (:641-643)
        return t.view(B, L, H, -1).permute(0, 2, 1, 3)

    q, k, v = heads(lin("node_to_q", node)), heads(lin("node_to_k", node)), heads(lin("node_to_v", node))   # :638-640
    d = q.shape[-1]
    e = lin("edge_emb", edge, bias=False).view(B, L, L, H, d).permute(0, 3, 1, 2, 4)          # b h i j d   (:645-646)
    logit = torch.einsum("bhid,bhjd->bhij", q, k) + torch.einsum("bhid,bhijd->bhij", q, e)    # :648-649
    att = logit * d ** (-0.5)                                                                 # :651, scale :616
    if mask is not None:
        att = att + ((1.0 - mask) * (-1e9))[:, None]                                          # :653-656
    att = att.softmax(dim=-1)                                                                 # :658
    upd = torch.einsum("bhij,bhjd->bhid", att, v) + torch.einsum("bhij,bhijd->bhid", att, e)  # :661-662
    upd = upd.permute(0, 2, 1, 3).reshape(B, L, H * d)                                        # :663
    return lin("node_update", node) + upd                                                     # :664


def graph_transformer_block(node, edge, mask, w, n_heads):
    """:676-677: to_out(ln(attn(...))) + node, to_out = Linear + ELU."""
    sub = {k[len("attn."):]: v for k, v in w.items() if k.startswith("attn.")}
    a = graph_transformer(node, edge, mask, sub, n_heads)
    a = F.layer_norm(a, (a.shape[-1],), w["ln.weight"], w["ln.bias"], 1e-5)
    return F.elu(F.linear(a, w["to_out.0.weight"], w["to_out.0.bias"])) + node

"""ORACLE shim: dgl.nn.pytorch.softmax.edge_softmax - softmax of edge scores over the edges that share a
destination node (the reference's call: equivariant_attention/modules.py GMABSE3)."""
import torch


def edge_softmax(graph, logits, eids=None, norm_by="dst"):
    assert norm_by == "dst" and eids is None
    idx, n = graph._dst, graph.num_nodes()
    shape = (n,) + tuple(logits.shape[1:])
    gather = idx.view((-1,) + (1,) * (logits.dim() - 1)).expand_as(logits)
    mx = torch.full(shape, float("-inf"), dtype=logits.dtype, device=logits.device)
    mx = mx.scatter_reduce(0, gather, logits, reduce="amax", include_self=True)
    e = torch.exp(logits - mx[idx])
    den = torch.zeros(shape, dtype=logits.dtype, device=logits.device).index_add_(0, idx, e)
    return e / den[idx]

"""ORACLE shim: a small functional stand-in for `dgl` (an un-vendored dependency of the reference's SE(3)
structure track, absent from this image). It implements exactly the message-passing surface the reference
touches (rosettafold_pytorch.py:856-860, equivariant_attention/modules.py) in plain PyTorch, so the WHOLE
reference model can run on the CPU in the build container. The SE(3) track is outside the hot path and is
shared by the reference and the accelerated model, so tests that compare the two are insensitive to this shim;
its arithmetic is nevertheless the documented DGL semantics (PARITY UNPINNED: no real dgl to compare with)."""
import contextlib

import torch

from . import function  # noqa: F401

__version__ = "1.1.0"


class _EdgeBatch:
    def __init__(self, g):
        self.src = {k: v[g._src] for k, v in g.ndata.items()}
        self.dst = {k: v[g._dst] for k, v in g.ndata.items()}
        self.data = g.edata


class DGLGraph:
    def __init__(self, src, dst, num_nodes):
        self._src, self._dst, self._n = src.long(), dst.long(), int(num_nodes)
        self.ndata, self.edata = {}, {}

    def to(self, device):
        self._src, self._dst = self._src.to(device), self._dst.to(device)
        self.ndata = {k: v.to(device) for k, v in self.ndata.items()}
        self.edata = {k: v.to(device) for k, v in self.edata.items()}
        return self

    @property
    def device(self):
        return self._src.device

    def num_nodes(self):
        return self._n

    number_of_nodes = num_nodes

    def num_edges(self):
        return int(self._src.numel())

    number_of_edges = num_edges

    def all_edges(self, form="uv", order=None):
        return self._src, self._dst

    edges = all_edges

    @contextlib.contextmanager
    def local_scope(self):
        nd, ed = dict(self.ndata), dict(self.edata)
        try:
            yield
        finally:
            self.ndata, self.edata = nd, ed

    def apply_edges(self, func):
        if isinstance(func, function._Builtin):
            assert func.kind == "e_dot_v"
            e, v = self.edata[func.a], self.ndata[func.b][self._dst]
            self.edata[func.out] = (e * v).sum(-1, keepdim=True)
        else:
            self.edata.update(func(_EdgeBatch(self)))

    def update_all(self, message_func, reduce_func):
        msgs = message_func(_EdgeBatch(self))
        assert isinstance(reduce_func, function._Builtin) and reduce_func.kind in ("mean", "sum")
        m = msgs[reduce_func.a]
        out = torch.zeros((self._n,) + tuple(m.shape[1:]), dtype=m.dtype, device=m.device)
        out.index_add_(0, self._dst, m)
        if reduce_func.kind == "mean":
            deg = torch.zeros(self._n, dtype=m.dtype, device=m.device)
            deg.index_add_(0, self._dst, torch.ones_like(self._dst, dtype=m.dtype))
            out = out / deg.clamp_min(1).view((-1,) + (1,) * (m.dim() - 1))
        self.ndata[reduce_func.out] = out


def graph(data, num_nodes=None, **_):
    src, dst = data
    return DGLGraph(torch.as_tensor(src), torch.as_tensor(dst), num_nodes if num_nodes is not None else int(max(src.max(), dst.max())) + 1)

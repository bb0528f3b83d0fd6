"""Dense graph transformer of the reference's initial-coordinate generator:
`GraphTransformer` (rosettafold_pytorch.py:613-664) and `GraphTransformerBlock` (:667-677) as drop-in `nn.Module`s with
the reference's constructor signatures and state_dict keys, on librfk.

The reference materialises the per-edge embedding e = edge_emb(edge_feat) as a [B, L, L, H*d] tensor (268 MB at L = 512
with the default H*d = 256) and contracts it twice with einsums. Both contractions are linear in e, so e never has to
exist:
    logit_e[b,h,i,j] = sum_d q[b,i,h,d] e[b,i,j,h,d] = sum_c edge[b,i,j,c] * qW[b,i,h,c],   qW = q . W_e   (per head)
    upd_e[b,i,h,d]   = sum_j att[b,h,i,j] e[b,i,j,h,d] = sum_c t[b,i,h,c] W_e[(h,d),c],    t = att . edge (per i)
i.e. four batched GEMMs over the d_edge-channel edge features themselves. Everything is rfk_gemm (strided batched views,
fused bias / scale / residual epilogues), rfk_softmax_rows, rfk_layernorm and rfk_convert_rows; the 1/sqrt(d) scale is
folded into the q projection (the logits are linear in q). Eval mode (attention dropout = identity), no autograd.
"""
import torch
from torch import nn

from . import modules as M
from . import ops
from .ops import cview


def _up8(n):
    return (n + 7) // 8 * 8


class GraphTransformer(nn.Module):
    """:613-664. forward(node_feat [B,L,Dn], edge_feat [B,L,L,De], edge_mask [B,L,L] or None) -> [B, L, H*d]."""

    def __init__(self, d_node_in, d_node_out, d_edge, n_heads, p_dropout=0.15):
        super().__init__()
        self.scale = d_node_out ** (-0.5)
        self.node_update = nn.Linear(d_node_in, d_node_out * n_heads, bias=True)
        self.node_to_q = nn.Linear(d_node_in, d_node_out * n_heads, bias=True)
        self.node_to_k = nn.Linear(d_node_in, d_node_out * n_heads, bias=True)
        self.node_to_v = nn.Linear(d_node_in, d_node_out * n_heads, bias=True)
        self.edge_emb = nn.Linear(d_edge, d_node_out * n_heads, bias=False)
        self.att_dropout = nn.Dropout(p_dropout)
        self.n_heads = n_heads
        self.d_in, self.d_out, self.d_edge = d_node_in, d_node_out, d_edge
        if d_node_in % 8 or d_node_out % 8 or d_edge % 8:
            raise NotImplementedError("GraphTransformer: feature widths must be multiples of 8 (16-byte operand rows)")

    def _pack(self):
        def build():
            dt = M._bdt()
            H, d, De = self.n_heads, self.d_out, self.d_edge
            We = self.edge_emb.weight.detach().view(H, d, De)
            return dict(
                Wq=M._w(self.node_to_q.weight * self.scale, dt), bq=M._f(self.node_to_q.bias * self.scale),
                Wk=M._w(self.node_to_k.weight, dt), bk=M._f(self.node_to_k.bias),
                Wv=M._w(self.node_to_v.weight, dt), bv=M._f(self.node_to_v.bias),
                Wu=M._w(self.node_update.weight, dt), bu=M._f(self.node_update.bias),
                We=We.to(dt).contiguous(),                   # [H, d, De]: B operand (n = d, k = c) of upd_e
                WeT=We.transpose(1, 2).to(dt).contiguous())  # [H, De, d]: B operand (n = c, k = d) of qW
        return M._packed(self, build)

    @torch.no_grad()
    def forward(self, node_feat, edge_feat, edge_mask=None):
        node = M._as_f32(node_feat).contiguous()
        edge = M._as_f32(edge_feat).contiguous()
        B, L, Dn = node.shape
        H, d, De = self.n_heads, self.d_out, self.d_edge
        if tuple(edge.shape) != (B, L, L, De) or Dn != self.d_in:
            raise ValueError("GraphTransformer: node_feat [B,L,d_node_in], edge_feat [B,L,L,d_edge] expected")
        pk = self._pack()
        dt = M._bdt()  # one operand dtype throughout: node / edge features follow an ELU or a LayerNorm (range-bounded)
        T, Lp = B * L, _up8(L)
        x = node.view(T, Dn) if M._MODE == 1 else ops.convert_rows(node.view(T, Dn), M._empty((T, Dn), dt, node))
        e_op = edge if M._MODE == 1 else ops.convert_rows(edge.view(-1, De), M._empty((B * L * L, De), dt, node)).view(B, L, L, De)
        # edge features with j contiguous: the K-major B operand of t = att . edge (the one copy that is a transposition)
        eT = torch.zeros((B, L, De, Lp), dtype=dt, device=node.device)
        eT[..., :L].copy_(e_op.transpose(2, 3))
        # projections, written in the layouts the contractions want (no rearrange copies):
        q = M._empty((B, L, H, d), dt, node)                 # scale folded in
        k = M._empty((B, H, L, d), dt, node)
        vT = torch.zeros((B, H, d, Lp), dtype=dt, device=node.device)
        u = M._empty((B, L, H * d), torch.float32, node)
        ops.gemm(x, pk["Wq"], cview(q.view(T, H * d)), bias=pk["bq"])
        ops.gemm(x.view(B, L, Dn), pk["Wk"][None], _kview(k), bias=pk["bk"])
        ops.gemm(x.view(B, L, Dn), pk["Wv"][None], _vview(vT, L), bias=pk["bv"])
        ops.gemm(x, pk["Wu"], cview(u.view(T, H * d)), bias=pk["bu"])
        # qW[b,i,h,c] = sum_d q[b,i,h,d] W_e[(h,d),c]          (batch = head)
        qW = M._empty((B, L, H, De), dt, node)
        ops.gemm(q.view(T, H, d).permute(1, 0, 2), pk["WeT"], _hview(qW.view(T, H, De)))
        # logits[b,i,h,j] = q.k + edge.qW (+ mask), j contiguous for the softmax
        logits = M._empty((B, L, H, Lp), torch.float32, node)
        lg = logits[..., :L]
        #   q.k: batch (b, h), m = i, n = j
        ops.gemm(q.permute(0, 2, 1, 3), k, _bh_view(lg))
        #   edge.qW: batch (b, i), m = j, n = h; accumulates onto q.k through the residual input (+ the additive mask)
        acc = _bi_view(lg)
        mask = None
        if edge_mask is not None:
            add = ((1.0 - M._as_f32(edge_mask)) * (-1e9)).contiguous()                    # [B, L(i), L(j)]   (:645-647)
            mask = add.view(1, B, L, 1, L, 1, 1).expand(1, B, L, 1, L, 1, H)               # broadcast over heads
        ops.gemm(e_op, qW, acc, r0=acc, r1=mask)
        att = torch.zeros((B, L, H, Lp), dtype=dt, device=node.device)
        ops.softmax_rows(logits.view(T * H, Lp)[:, :L], att.view(T * H, Lp)[:, :L])
        # updated[b,i,h,d] = att.v + (att.edge).W_e
        upd = M._empty((B, L, H, d), torch.float32, node)
        ops.gemm(att[..., :L].permute(0, 2, 1, 3), vT[..., :L], _bh_out(upd), r0=_bh_out(u.view(B, L, H, d)))
        t = M._empty((B, L, H, De), dt, node)
        ops.gemm(att[..., :L], eT[..., :L], t.view(1, B, L, 1, H, 1, De))
        ops.gemm(t.view(T, H, De).permute(1, 0, 2), pk["We"], _hview(upd.view(T, H, d)), r0=_hview(upd.view(T, H, d)))
        return upd.view(B, L, H * d)


def _hview(t3):
    """[T, H, n] buffer as the c_view of a GEMM batched over heads (Z0 = h, m = token, n = last dim)."""
    T, H, n = t3.shape
    return t3.as_strided((1, 1, H, T, 1, 1, n), (0, 0, t3.stride(1), t3.stride(0), 0, 0, t3.stride(2)))


def _kview(k):
    """k buffer [B, H, L, d] as the output of x[b] . Wk^T with n = (h, d): batch b, m = l, n split (N1 = h, NR = d)."""
    B, H, L, d = k.shape
    return k.as_strided((1, 1, B, L, 1, H, d), (0, 0, k.stride(0), k.stride(2), 0, k.stride(1), k.stride(3)))


def _vview(vT, L):
    """vT buffer [B, H, d, Lp] as the output of x[b] . Wv^T: m = l (stride 1), n split (N1 = h, NR = d)."""
    B, H, d, _ = vT.shape
    return vT.as_strided((1, 1, B, L, 1, H, d), (0, 0, vT.stride(0), vT.stride(3), 0, vT.stride(1), vT.stride(2)))


def _bh_view(lg):
    """logits[b, i, h, j] as the output of a GEMM batched over (b, h): m = i, n = j."""
    B, L, H, Lj = lg.shape
    return lg.as_strided((1, B, H, L, 1, 1, Lj), (0, lg.stride(0), lg.stride(2), lg.stride(1), 0, 0, lg.stride(3)))


def _bi_view(lg):
    """logits[b, i, h, j] as the output of a GEMM batched over (b, i): m = j, n = h."""
    B, L, H, Lj = lg.shape
    return lg.as_strided((1, B, L, 1, Lj, 1, H), (0, lg.stride(0), lg.stride(1), 0, lg.stride(3), 0, lg.stride(2)))


def _bh_out(t4):
    """[B, L, H, d] buffer as the output / residual of a GEMM batched over (b, h): m = i, n = d."""
    B, L, H, d = t4.shape
    return t4.as_strided((1, B, H, L, 1, 1, d), (0, t4.stride(0), t4.stride(2), t4.stride(1), 0, 0, t4.stride(3)))


class GraphTransformerBlock(nn.Module):
    """:667-677: to_out(LN(attn(node, edge, mask))) + node with to_out = Linear + ELU."""

    def __init__(self, d_node_in, d_node_out, d_edge, n_heads, p_dropout=0.15):
        super().__init__()
        self.attn = GraphTransformer(d_node_in, d_node_out, d_edge, n_heads, p_dropout)
        self.ln = nn.LayerNorm(d_node_out * n_heads)
        self.to_out = nn.Sequential(nn.Linear(d_node_out * n_heads, d_node_in), nn.ELU())

    def _pack(self):
        def build():
            return dict(W=M._w(self.to_out[0].weight, M._bdt()), b=M._f(self.to_out[0].bias))
        return M._packed(self.to_out, build)

    @torch.no_grad()
    def forward(self, node_feat, edge_feat, edge_mask=None):
        node = M._as_f32(node_feat).contiguous()
        B, L, Dn = node.shape
        a = self.attn(node, edge_feat, edge_mask)                                  # f32 [B, L, H*d]
        T, Dh = B * L, a.shape[-1]
        an = M._ln_into(a.view(T, Dh), self.ln, M._empty((T, Dh), M._bdt(), node))
        pk = self._pack()
        out = M._empty((T, Dn), torch.float32, node)
        # ELU is applied before the residual add by the epilogue: elu(x W^T + b) + node   (:677)
        ops.gemm(an, pk["W"], cview(out), bias=pk["b"], act=ops.ACT_ELU, r0=cview(node.view(T, Dn)))
        return out.view(B, L, Dn)

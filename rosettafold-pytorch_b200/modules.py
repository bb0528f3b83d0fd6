"""Drop-in nn.Module mirrors of the reference trunk (rosettafold_pytorch/rosettafold_pytorch.py).

Same class names, constructor signatures, forward signatures/returns and parameter names
(state_dict keys) as the reference lines cited on each class, so reference weights load
one-to-one. The forwards are inference (eval-mode) computations built ONLY from librfk kernels
(`ops`): dropout is the identity and autograd is not recorded. torch is used for allocation and
views. The 3x3 Conv2d pair of PairUpdateWithMsa (:451-457) runs on the tcgen05 implicit-GEMM kernel
in bf16 mode and on a SIMT fp32 kernel in the fp32 validation mode: no library kernel is on the path.

Residual streams (msa, pair) are float32. `set_mode("bf16")` (default) feeds bf16 operands to the
tcgen05 kernels with fp32 accumulation; `set_mode("fp32")` is the fp32 validation mode.

Deliberate deviations from the reference, all documented in DESIGN.md:
  * MsaUpdateWithPair.encoder_layers is an nn.ModuleList (the reference's plain list, :602-605,
    hides its weights from .eval()/.to()/state_dict()); `load_reference_weights` copies them.
  * every module works on the device of its input (the reference's embeddings pin CPU tensors).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .ops import ACT_RELU, EPI_BLOCKLN32, cview

_MODE = 0  # 0 = 16-bit tensor-core mode ("bf16"), 1 = fp32 validation mode
# 16-bit format of the RANGE-BOUNDED operands in the tensor-core mode: the whole MSA track (LayerNorm outputs, the
# projections q / k / v / FeedForward hidden of those, softmax probabilities, the softmax-kernel FAVOR features) and the
# LayerNorm-ed operands in front of the convolution block of PairUpdateWithMsa. IEEE half runs at the bf16 tensor-core
# rate with 11 instead of 8 significand bits: it cuts the rounding error of those stages ~8x, which is what keeps a
# 13-block trunk inside the 1e-2 budget (profiles/r02_parity.md). Every such value is bounded by
# sqrt(d) * max|gamma| * max row-norm of the weights, far below half's 65504 for any sane weights;
# set_bounded_operand_dtype("bf16") restores bf16 everywhere. The unbounded pair-track operands (ReLU-kernel FAVOR
# features and contexts, which sum over up to thousands of tokens, pair FeedForward hidden, convolution
# activations) always stay bf16.
_BOUNDED_DT = torch.float16


def set_mode(mode: str):
    global _MODE
    if mode not in ("bf16", "fp32"):
        raise ValueError("mode must be 'bf16' or 'fp32'")
    _MODE = 0 if mode == "bf16" else 1


def get_mode() -> str:
    return "bf16" if _MODE == 0 else "fp32"


def set_bounded_operand_dtype(name: str):
    """"f16" (default) or "bf16": the 16-bit format of the range-bounded operands in the tensor-core mode."""
    global _BOUNDED_DT
    if name not in ("f16", "bf16"):
        raise ValueError("bounded operand dtype must be 'f16' or 'bf16'")
    _BOUNDED_DT = torch.float16 if name == "f16" else torch.bfloat16


def _adt():
    """Operand dtype of the unbounded pair-track tensors."""
    return torch.bfloat16 if _MODE == 0 else torch.float32


def _bdt():
    """Operand dtype of the range-bounded tensors (see _BOUNDED_DT)."""
    return _BOUNDED_DT if _MODE == 0 else torch.float32


def _up8(n: int) -> int:
    return (n + 7) // 8 * 8


# ---------------------------------------------------------------------------------------------
# weight packing cache: derived tensors (dtype casts, concatenations, folded affines) are built
# once per (mode, device) and rebuilt when any parameter is modified in place or replaced.
# ---------------------------------------------------------------------------------------------
def _packed(module: nn.Module, builder):
    params = list(module.parameters()) + list(module.buffers())
    sig = (_MODE, _BOUNDED_DT, tuple((p.data_ptr(), p._version, p.device) for p in params))
    cache = module.__dict__.get("_rfk_pack")
    if cache is None or cache[0] != sig:
        with torch.no_grad():
            cache = (sig, builder())
        module.__dict__["_rfk_pack"] = cache
    return cache[1]


def _w(t: torch.Tensor, dtype) -> torch.Tensor:
    """Weight matrix [N, K] in the operand dtype `dtype` with K padded to a multiple of 8 elements."""
    N, K = t.shape
    buf = torch.zeros((N, _up8(K)), dtype=dtype, device=t.device)
    buf[:, :K] = t.detach().to(dtype)
    return buf[:, :K]


def _f(t):
    return None if t is None else t.detach().float().contiguous()


def _empty(shape, dtype, like):
    return torch.empty(shape, dtype=dtype, device=like.device)


def _ln_into(x2, ln: nn.LayerNorm, out):
    ops.layernorm(x2, _f(ln.weight), _f(ln.bias), ln.eps, out)
    return out


def _as_f32(x):
    return x if x.dtype == torch.float32 else x.float()


# ---------------------------------------------------------------------------------------------
# generic containers (API parity; the fused forwards below never route through these)
# ---------------------------------------------------------------------------------------------
class Residual(nn.Module):
    """reference :18-28"""

    def __init__(self, fn, p_dropout=None):
        super().__init__()
        self.fn = fn
        self.dropout = nn.Dropout(p_dropout) if p_dropout is not None else None

    def forward(self, x):
        # generic container form only: every Residual the trunk builds is executed by its owner's fused path (the add
        # lives in a GEMM / LayerNorm / InstanceNorm epilogue); this plain add serves a user-supplied `fn` and is
        # eval-mode like the rest of the package (the dropout of :27 is the identity)
        return self.fn(x) + x


class ColWise(nn.Module):
    """reference :31-41 — fn attends over axis 2 of (b, n, l, d)."""

    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def forward(self, x):
        if isinstance(self.fn, PerformerSelfAttention):
            return self.fn._attend(x, token_dim=2)
        b, n = x.shape[:2]
        return self.fn(x.reshape(b * n, *x.shape[2:])).reshape(x.shape)


class RowWise(nn.Module):
    """reference :44-54 — fn attends over axis 1 of (b, n, l, d)."""

    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def forward(self, x):
        if isinstance(self.fn, PerformerSelfAttention):
            return self.fn._attend(x, token_dim=1)
        b, n, l = x.shape[:3]
        y = self.fn(x.transpose(1, 2).reshape(b * l, n, -1))
        return y.reshape(b, l, n, -1).transpose(1, 2)


# ---------------------------------------------------------------------------------------------
# FeedForward (:270-281)
# ---------------------------------------------------------------------------------------------
class FeedForward(nn.Module):
    def __init__(self, d_emb, d_ff, p_dropout=0.1):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(d_emb, d_ff), nn.ReLU(), nn.Dropout(p_dropout),
                                 nn.Linear(d_ff, d_emb))

    def _pack(self):
        def build():
            d = dict(b1=_f(self.net[0].bias), b2=_f(self.net[3].bias))
            for dt in {_adt(), _bdt()}:  # pair-track and MSA-track callers
                d[dt] = (_w(self.net[0].weight, dt), _w(self.net[3].weight, dt))
            return d
        return _packed(self, build)

    def _run(self, xn2, res2=None, out_dtype=torch.float32):
        """xn2: [T, d_emb] in an operand dtype (it selects the format of the weights and of the hidden tensor);
        returns W2 relu(W1 xn + b1) + b2 (+ res2)."""
        pk = self._pack()
        W1, W2 = pk[xn2.dtype]
        T = xn2.shape[0]
        hid = _empty((T, W1.shape[0]), xn2.dtype, xn2)
        ops.gemm(xn2, W1, cview(hid), bias=pk["b1"], act=ACT_RELU)
        out = _empty((T, W2.shape[0]), out_dtype, xn2)
        ops.gemm(hid, W2, cview(out), bias=pk["b2"],
                 r0=None if res2 is None else cview(res2))
        return out

    @torch.no_grad()
    def forward(self, x):
        x2 = _as_f32(x).reshape(-1, x.shape[-1])
        xin = x2 if _MODE == 1 else ops.convert_rows(x2, _empty(x2.shape, _adt(), x2))
        return self._run(xin).reshape(x.shape)


def _ff_block(ln: nn.LayerNorm, ff: FeedForward, x2: torch.Tensor, dt) -> torch.Tensor:
    """x + FF(LN(x)) on a [T, D] float32 residual stream (:326-332, :352); dt: operand dtype of the track."""
    xn = _ln_into(x2, ln, _empty(x2.shape, dt, x2))
    return ff._run(xn, res2=x2)


# ---------------------------------------------------------------------------------------------
# PositionWiseWeightFactor (:184-217)
# ---------------------------------------------------------------------------------------------
class PositionWiseWeightFactor(nn.Module):
    def __init__(self, d_msa=384, n_heads=12, p_dropout=0.1):
        super().__init__()
        assert (
            d_msa % n_heads == 0
        ), f"[{self.__class__.__name__}]: d_msa ({d_msa}) must be divisible by n_heads ({n_heads})."
        self.n_heads = n_heads
        self.d_head = d_msa // n_heads
        self.scale = self.d_head ** (-0.5)
        self.to_q = nn.Sequential(nn.Linear(d_msa, d_msa))
        self.to_k = nn.Sequential(nn.Linear(d_msa, d_msa))
        self.dropout = nn.Dropout(p_dropout)

    def _pack(self):
        dt = _bdt()  # every caller hands LayerNorm outputs: range-bounded operands
        return _packed(self, lambda: dict(Wq=_w(self.to_q[0].weight, dt), bq=_f(self.to_q[0].bias),
                                          Wk=_w(self.to_k[0].weight, dt), bk=_f(self.to_k[0].bias)))

    def _project_query(self, x4):
        """x4: [B,N,L,D] operand dtype -> pq [B,L,D] = to_q(x[:, 0]) (:207-209, unscaled)."""
        pk = self._pack()
        B, N, L, D = x4.shape
        pq = _empty((B, L, D), x4.dtype, x4)
        ops.gemm(x4[:, 0], pk["Wq"], pq.view(1, 1, B, 1, L, 1, D), bias=pk["bq"])
        return pq

    def _weights(self, x4):
        """x4: [B,N,L,D] operand dtype -> w [B,N,L,H] float32."""
        pk = self._pack()
        B, N, L, D = x4.shape
        pq = self._project_query(x4)
        pkk = _empty((B, N, L, D), x4.dtype, x4)
        ops.gemm(x4.reshape(-1, D), pk["Wk"], cview(pkk.view(-1, D)), bias=pk["bk"])
        w = _empty((B, N, L, self.n_heads), torch.float32, x4)
        ops.poswise_weight(pq, pkk, self.scale, w_out=w, heads=self.n_heads, d_head=self.d_head)
        return w

    @torch.no_grad()
    def forward(self, msa_emb):
        """msa : (B, N, L, d_msa) -> (B, N, h, L, 1)"""
        x = _as_f32(msa_emb).contiguous()
        xin = x if _MODE == 1 else ops.convert_rows(x.view(-1, x.shape[-1]),
                                                    _empty((x.numel() // x.shape[-1], x.shape[-1]), _bdt(), x)).view(x.shape)
        w = self._weights(xin)
        return w.permute(0, 1, 3, 2).unsqueeze(-1)


# ---------------------------------------------------------------------------------------------
# SoftTiedAttentionOverResidues (:220-267)
# ---------------------------------------------------------------------------------------------
class SoftTiedAttentionOverResidues(nn.Module):
    def __init__(self, d_msa=384, n_heads=12, p_dropout=0.1, return_att=False):
        super().__init__()
        assert (
            d_msa % n_heads == 0
        ), f"[{self.__class__.__name__}]: d_msa ({d_msa}) must be divisible by n_heads ({n_heads})."
        self.n_heads = n_heads
        self.d_head = d_msa // n_heads
        self.scale = self.d_head ** (-0.5)
        self.return_att = return_att
        self.poswise_weight = PositionWiseWeightFactor(d_msa, n_heads, p_dropout)
        self.to_q = nn.Linear(d_msa, d_msa)
        self.to_k = nn.Linear(d_msa, d_msa)
        self.to_v = nn.Linear(d_msa, d_msa)
        self.to_out = nn.Linear(d_msa, d_msa)
        self.dropout = nn.Dropout(p_dropout)

    def _pack(self):
        def build():
            pw, dt = self.poswise_weight, _bdt()
            return dict(
                # fused [q | poswise-k] projection: both stay in the (b n l) row layout
                Wqp=_w(torch.cat([self.to_q.weight, pw.to_k[0].weight], 0), dt),
                bqp=_f(torch.cat([self.to_q.bias, pw.to_k[0].bias], 0)),
                Wk=_w(self.to_k.weight, dt), bk=_f(self.to_k.bias),
                Wv=_w(self.to_v.weight, dt), bv=_f(self.to_v.bias),
                Wo=_w(self.to_out.weight, dt), bo=_f(self.to_out.bias))
        return _packed(self, build)

    def _attend(self, xn4, res2, want_att, shard=None):
        """xn4: [B,N,L,D] operand dtype (already normalised by the caller).
        Returns (to_out(attention) (+ res2) as float32 [T, D], symmetrised att or None).

        With `shard` (rosettafold_pytorch_b200.sharded.SequenceShard) xn4 holds this rank's slice of the
        sequences: projections, q scaling, A.V and to_out are per sequence and stay local; the three places
        where :205-257 couple the sequences go through the context - the query row (sequence 0 of the whole
        MSA, :207), the softmax over ALL sequences of the position-wise weights (:213: merged from per-shard
        (max, sum) statistics) and the logits, which sum over all sequences (:254: all-reduce)."""
        pk = self._pack()
        B, N, L, D = xn4.shape
        H, dh = self.n_heads, self.d_head
        T = B * N * L
        adt = xn4.dtype
        xn2 = xn4.view(T, D)
        xb = xn4.view(B, N * L, D)
        # q and the poswise keys (plain rows); K and V go straight to their contraction layouts
        qp = _empty((T, 2 * D), adt, xn4)
        ops.gemm(xn2, pk["Wqp"], cview(qp), bias=pk["bqp"])
        kt = _empty((B, H, L, N * dh), adt, xn4)  # b h j (n d): K-major operand of the logits
        ops.gemm(xb, pk["Wk"][None], kt.view(B, H, L, N, dh).permute(0, 3, 2, 1, 4)[None, None],
                 bias=pk["bk"])
        Lp = _up8(L)
        vt = _empty((B, H, N * dh, Lp), adt, xn4)  # b h (n d) j: K-major operand of A.V
        ops.gemm(xb, pk["Wv"][None],
                 vt.view(B, H, N, dh, Lp)[..., :L].permute(0, 2, 4, 1, 3)[None, None], bias=pk["bv"])
        pq = self.poswise_weight._project_query(xn4 if shard is None else shard.first_sequence(xn4))
        qp4 = qp.view(B, N, L, 2 * D)
        qt = _empty((B, H, L, N * dh), adt, xn4)  # q * w * scale, b h i (n d)
        stats = None if shard is None else _empty((B, L, H, 2), torch.float32, xn4)
        ops.poswise_weight(pq, qp4[..., D:], self.poswise_weight.scale, q=qp4[..., :D],
                           q_scale=self.scale, qt=qt, heads=H, d_head=dh, stats=stats)
        logits = _empty((B, H, L, L), torch.float32, xn4)
        ops.gemm(qt, kt, logits.view(1, B, H, 1, L, 1, L))
        A = _empty((B, H, L, Lp), adt, xn4)
        R = B * H * L  # rows of the attention map
        if shard is not None:
            # this shard's weights were normalised over its own sequences: rescale row (h, i) of the partial
            # logits to the global normalisation, then sum the partial logits over the shards
            logits.mul_(shard.softmax_correction(stats).permute(0, 2, 1).unsqueeze(-1))
        if shard is not None and R % shard.world == 0:
            # reduce-scatter by rows -> softmax of this rank's rows -> all-gather of the 16-bit probabilities: three
            # quarters of the bytes of an all-reduce of the fp32 map, and 1 / P of the softmax
            mine = shard.reduce_scatter_rows(logits.view(R, L))
            A_mine = torch.zeros((mine.shape[0], Lp), dtype=adt, device=xn4.device)
            ops.softmax_rows(mine, A_mine[:, :L])
            shard.all_gather_rows(A_mine, A.view(R, Lp))
        else:
            if shard is not None:
                shard.allreduce(logits)
            ops.softmax_rows(logits.view(R, L), A.view(R, Lp)[:, :L])
        att = None
        if want_att:
            att = _empty((B, L, L, H), torch.float32, xn4)
            ops.tied_att_symmetrize(A[..., :L], att)
        o = _empty((B, N, L, D), adt, xn4)
        ops.gemm(A[..., :L], vt[..., :L], o.view(B, N, L, H, dh).permute(0, 3, 2, 1, 4).unsqueeze(2)[None])
        out = _empty((T, D), torch.float32, xn4)
        ops.gemm(o.view(T, D), pk["Wo"], cview(out), bias=pk["bo"],
                 r0=None if res2 is None else cview(res2))
        return out, att

    @torch.no_grad()
    def forward(self, x):
        """x : (B, N, L, d_msa)"""
        x = _as_f32(x).contiguous()
        D = x.shape[-1]
        xin = x if _MODE == 1 else ops.convert_rows(x.view(-1, D), _empty((x.numel() // D, D), _bdt(), x)).view(x.shape)
        out, att = self._attend(xin, None, self.return_att)
        out = out.view(x.shape)
        return (out, att) if self.return_att else out


# ---------------------------------------------------------------------------------------------
# Performer SelfAttention (performer_pytorch.SelfAttention as called at :313-318, :505-518)
# ---------------------------------------------------------------------------------------------
class _FastAttention(nn.Module):
    def __init__(self, dim_heads, nb_features, generalized_attention):
        super().__init__()
        from ._favor_init import gaussian_orthogonal_random_matrix

        self.nb_features = nb_features
        self.generalized_attention = generalized_attention
        self.register_buffer("projection_matrix",
                             gaussian_orthogonal_random_matrix(nb_features, dim_heads))


class PerformerSelfAttention(nn.Module):
    """FAVOR+ self-attention with the constructor surface the reference uses
    (dim, heads, dropout, generalized_attention; performer_kws must be empty)."""

    DIM_HEAD = 64

    def __init__(self, dim, heads=8, dropout=0.0, generalized_attention=False, **performer_kws):
        super().__init__()
        if performer_kws:
            raise NotImplementedError(f"unsupported performer_kws: {sorted(performer_kws)}")
        import math

        inner = self.DIM_HEAD * heads
        self.heads = heads
        self.fast_attention = _FastAttention(self.DIM_HEAD, int(self.DIM_HEAD * math.log(self.DIM_HEAD)),
                                             generalized_attention)
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_k = nn.Linear(dim, inner, bias=False)
        self.to_v = nn.Linear(dim, inner, bias=False)
        self.to_out = nn.Linear(inner, dim, bias=True)
        self.dropout = nn.Dropout(dropout)

    def _dt(self):
        """Operand dtype: the softmax kernel (MSA columns) works on bounded features, the ReLU kernel (pair axes) does not."""
        return _adt() if self.fast_attention.generalized_attention else _bdt()

    def _pack(self):
        dt = self._dt()
        return _packed(self, lambda: dict(
            Wqkv=_w(torch.cat([self.to_q.weight, self.to_k.weight, self.to_v.weight], 0), dt),
            Wo=_w(self.to_out.weight, dt), bo=_f(self.to_out.bias),
            proj=_f(self.fast_attention.projection_matrix)))

    def _run(self, xn4, token_dim, res2, out_dtype=torch.float32):
        """xn4: [A0,A1,A2,D] operand dtype; attention over axis `token_dim` (1 or 2), batched over
        the other two axes. Returns to_out(attn) (+ res2) as `out_dtype` [T, D]."""
        pk = self._pack()
        A0, A1, A2, D = xn4.shape
        T = A0 * A1 * A2
        inner = self.heads * self.DIM_HEAD
        adt = xn4.dtype
        qkv = _empty((T, 3 * inner), adt, xn4)
        ops.gemm(xn4.view(T, D), pk["Wqkv"], cview(qkv))
        ao = _empty((T, inner), adt, xn4)
        q4 = qkv.view(A0, A1, A2, 3 * inner)
        a4 = ao.view(A0, A1, A2, inner)
        if token_dim == 1:  # tokens along axis 1: group (a0, a2)
            q4, a4 = q4.permute(0, 2, 1, 3), a4.permute(0, 2, 1, 3)
        ops.favor_attention(q4[..., :inner], q4[..., inner:2 * inner], q4[..., 2 * inner:], a4,
                            pk["proj"], kind=1 if self.fast_attention.generalized_attention else 0,
                            heads=self.heads)
        out = _empty((T, D), out_dtype, xn4)
        ops.gemm(ao, pk["Wo"], cview(out), bias=pk["bo"],
                 r0=None if res2 is None else cview(res2))
        return out

    def _attend(self, x4, token_dim):
        x4 = _as_f32(x4).contiguous()
        D = x4.shape[-1]
        xin = x4 if _MODE == 1 else ops.convert_rows(x4.view(-1, D), _empty((x4.numel() // D, D), self._dt(), x4)).view(x4.shape)
        return self._run(xin, token_dim, None).view(x4.shape)

    @torch.no_grad()
    def forward(self, x):
        """x: (b, n, d) — attention over n (performer_pytorch.SelfAttention.forward)."""
        return self._attend(x.unsqueeze(0), token_dim=2).squeeze(0)


def _performer_block(ln: nn.LayerNorm, attn: PerformerSelfAttention, x4, token_dim):
    """x + attn(LN(x)) on a float32 [A0,A1,A2,D] stream, attention over `token_dim`."""
    D = x4.shape[-1]
    x2 = x4.view(-1, D)
    xn = _ln_into(x2, ln, _empty(x2.shape, attn._dt(), x2))
    return attn._run(xn.view(x4.shape), token_dim, x2).view(x4.shape)


# ---------------------------------------------------------------------------------------------
# EncoderLayer (:284-354) and MsaUpdateUsingSelfAttention (:357-409)
# ---------------------------------------------------------------------------------------------
class EncoderLayer(nn.Module):
    def __init__(self, d_msa=384, d_ff=384 * 4, n_heads=12, p_dropout=0.1, tied=False,
                 performer=False, performer_kws={}, return_att=False):
        super().__init__()
        self.tied = tied
        self.return_att = return_att
        if self.tied:
            self.attn = SoftTiedAttentionOverResidues(d_msa=d_msa, n_heads=n_heads,
                                                      p_dropout=p_dropout, return_att=return_att)
        elif performer:
            if return_att:
                raise NotImplementedError("PerformerSelfAttention does not support return_att.")
            self.attn = PerformerSelfAttention(dim=d_msa, heads=n_heads, dropout=p_dropout,
                                               **performer_kws)
        else:
            raise NotImplementedError
        self.ln = nn.LayerNorm(d_msa)
        self.dropout = nn.Dropout(p_dropout)
        self.ff = Residual(nn.Sequential(nn.LayerNorm(d_msa),
                                         FeedForward(d_msa, d_ff, p_dropout=p_dropout),
                                         nn.Dropout(p_dropout)))

    def _run(self, x4, token_dim=2, want_att=None, shard=None):
        """x4: float32 [B,N,L,D]. Tied: attention over residues with logits tied over N.
        Performer: attention over axis `token_dim`. `shard`: sequence-shard context (tied layers only,
        see SoftTiedAttentionOverResidues._attend)."""
        D = x4.shape[-1]
        x2 = x4.view(-1, D)
        att = None
        if self.tied:
            xn = _ln_into(x2, self.ln, _empty(x2.shape, _bdt(), x2))
            want = self.return_att if want_att is None else want_att
            x1, att = self.attn._attend(xn.view(x4.shape), x2, want, shard)
        else:
            x1 = _performer_block(self.ln, self.attn, x4, token_dim).view(-1, D)
        out = _ff_block(self.ff.fn[0], self.ff.fn[1], x1, _bdt()).view(x4.shape)
        return out, att

    @torch.no_grad()
    def forward(self, x):
        x = _as_f32(x).contiguous()
        out, att = self._run(x)
        return (out, att) if self.return_att else out


class MsaUpdateUsingSelfAttention(nn.Module):
    def __init__(self, d_msa=384, d_ff=384 * 4, n_heads=12, p_dropout=0.1, n_encoder_layers=4,
                 performer_kws={}):
        super().__init__()
        self.residue_wise_encoder_layers = nn.ModuleList([
            EncoderLayer(d_msa=d_msa, d_ff=d_ff, n_heads=n_heads, p_dropout=p_dropout, tied=True,
                         performer=False, return_att=True) for _ in range(n_encoder_layers)])
        self.sequence_wise_encoder_layers = nn.ModuleList([
            EncoderLayer(d_msa=d_msa, d_ff=d_ff, n_heads=n_heads, p_dropout=p_dropout, tied=False,
                         performer=True, performer_kws=performer_kws) for _ in range(n_encoder_layers)])

    @torch.no_grad()
    def forward(self, x):
        x = _as_f32(x).contiguous()
        att = None
        n = len(self.residue_wise_encoder_layers)
        for i, layer in enumerate(self.residue_wise_encoder_layers):
            # only the last layer's map is consumed (:400-401): skip the others' symmetrisation
            x, a = layer._run(x, want_att=(i == n - 1))
            att = a if a is not None else att
        # the reference transposes to (b l n d) here (:403); we attend over axis 1 in place
        for layer in self.sequence_wise_encoder_layers:
            x, _ = layer._run(x, token_dim=1)
        return x, att


# ---------------------------------------------------------------------------------------------
# OuterProductMean (:412-427) and PairUpdateWithMsa (:430-498)
# ---------------------------------------------------------------------------------------------
class OuterProductMean(nn.Module):
    def __init__(self, in_features, out_features):
        super().__init__()
        self.in_features = in_features
        self.to_out = nn.Sequential(nn.LayerNorm(in_features ** 2),
                                    nn.Linear(in_features ** 2, out_features))

    def _pack(self):
        def build():
            ln, lin = self.to_out[0], self.to_out[1]
            g, b = ln.weight.detach().float(), ln.bias.detach().float()
            Wl, bl = lin.weight.detach().float(), lin.bias.detach().float()
            # bf16 mode: the LayerNorm(1024) affine is folded into the Linear (exact algebra), so the
            # GEMM epilogue only normalises: Linear(g*xhat + b) = (W*g) xhat + (W b + bias)
            dt = _bdt()  # the operands are LayerNorm outputs (proj_msa, the in-epilogue LayerNorm(1024))
            return dict(g=g.contiguous(), b=b.contiguous(), W=_w(lin.weight, dt), bias=_f(lin.bias),
                        Wfold=_w(Wl * g[None, :], dt), bfold=(bl + Wl @ b).contiguous())
        return _packed(self, build)

    def _run(self, xt, yt, B, L):
        """xt: [B, Li*P, N], yt: [B, L*P, N] K-major operands (Li = L, or the rows of a row shard).
        Returns Linear(LN(outer-product sum)) f32 [B*Li*L, out]."""
        pk = self._pack()
        P = self.in_features
        adt = xt.dtype
        ln = self.to_out[0]
        Li = xt.shape[1] // P
        o = _empty((B, Li, L, P * P), adt, xt)
        ov = o.view(B, Li, L, P, P).permute(0, 1, 3, 2, 4)[None, None]  # [1,1,b,i,u,j,v]
        fused = _MODE == 0 and P == 32
        if fused:
            ops.gemm(xt, yt, ov, epi=EPI_BLOCKLN32, ln_eps=ln.eps)
        else:
            ops.gemm(xt, yt, ov)
            o2 = o.view(-1, P * P)
            ops.layernorm(o2, pk["g"], pk["b"], ln.eps, o2)
        out = _empty((B * Li * L, pk["W"].shape[0]), torch.float32, xt)
        ops.gemm(o.view(-1, P * P), pk["Wfold" if fused else "W"], cview(out),
                 bias=pk["bfold" if fused else "bias"])
        return out

    @torch.no_grad()
    def forward(self, x, y=None):
        y = x if y is None else y
        B, N, L, P = x.shape
        Np = _up8(N)
        adt = _bdt()
        xt = torch.zeros((B, L * P, Np), dtype=adt, device=x.device)
        yt = torch.zeros((B, L * P, Np), dtype=adt, device=x.device)
        # standalone API path only: relayout with torch (the fused path uses rfk_opm_prep)
        xt[..., :N] = x.permute(0, 2, 3, 1).reshape(B, L * P, N).to(adt)
        yt[..., :N] = y.permute(0, 2, 3, 1).reshape(B, L * P, N).to(adt)
        return self._run(xt[..., :N], yt[..., :N], B, L).view(B, L, L, -1)


class PairUpdateWithMsa(nn.Module):
    def __init__(self, d_msa, d_proj, d_pair, n_heads, p_dropout=0.1):
        super().__init__()
        self.d_proj, self.d_pair, self.n_heads = d_proj, d_pair, n_heads
        self.proj_msa = nn.Sequential(nn.LayerNorm(d_msa), nn.Linear(d_msa, d_proj), nn.LayerNorm(d_proj))
        self.poswise_weight = PositionWiseWeightFactor(d_proj, 1, p_dropout)
        self.outer_product_mean = OuterProductMean(d_proj, d_pair)
        self.ln_coevol_feat = nn.LayerNorm(d_pair)
        self.ln_pair = nn.LayerNorm(d_pair)
        d_feat_full = d_pair * 2 + d_proj * 4 + n_heads
        self.resnet = nn.Sequential(
            nn.Linear(d_feat_full, d_pair),
            Residual(nn.Sequential(
                nn.Identity(),  # Rearrange("b l1 l2 d -> b d l1 l2")
                nn.Conv2d(d_pair, d_pair, kernel_size=3, padding="same", bias=False),
                nn.InstanceNorm2d(d_pair, affine=True, eps=1e-6),
                nn.ELU(),
                nn.Dropout(p_dropout),
                nn.Conv2d(d_pair, d_pair, kernel_size=3, padding="same", bias=False),
                nn.InstanceNorm2d(d_pair, affine=True, eps=1e-6),
                nn.Identity(),  # Rearrange("b d l1 l2 -> b l1 l2 d")
            )),
            nn.ELU(),
        )

    def _pack(self):
        def build():
            P, Q, H = self.d_pair, self.d_proj, self.n_heads
            W = self.resnet[0].weight  # [d_pair, P | 2Q | 2Q | P | H]  (:487-496)
            c0, c1, c2, c3 = P, P + 2 * Q, P + 4 * Q, 2 * P + 4 * Q
            fn = self.resnet[1].fn
            bdt = _bdt()  # LayerNorm outputs and attention probabilities in front of the convolution block
            return dict(
                Wproj=_w(self.proj_msa[1].weight, bdt), bproj=_f(self.proj_msa[1].bias),
                # dense part of the 716-wide Linear: [coevol | ln_pair | att]
                Wf=_w(torch.cat([W[:, :c0], W[:, c2:c3], W[:, c3:]], 1), bdt), bf=_f(self.resnet[0].bias),
                # rank-1 parts: row-tiled and column-tiled msa_1d (fp32 SIMT GEMMs, K = 2Q)
                Wr=W[:, c0:c1].detach().float().contiguous(), Wc=W[:, c1:c2].detach().float().contiguous(),
                # bf16 mode: tap-major packed weights of the implicit-GEMM conv kernel; fp32
                # validation mode: plain fp32 weights for the library convolution
                conv1=ops.pack_conv3x3_weight(fn[1].weight) if _MODE == 0 else ops.pack_conv3x3_weight_f32(fn[1].weight),
                conv2=ops.pack_conv3x3_weight(fn[5].weight) if _MODE == 0 else ops.pack_conv3x3_weight_f32(fn[5].weight),
                g1=_f(fn[2].weight), b1=_f(fn[2].bias), g2=_f(fn[6].weight), b2=_f(fn[6].bias))
        return _packed(self, build)

    def _conv(self, x_bhwc, w):
        """3x3 'same' convolution on a channels-last [B,H,W,C] map: rfk_conv3x3_nhwc (tcgen05
        implicit GEMM) in bf16 mode, rfk_conv3x3_nhwc_f32 (SIMT fp32) in the fp32 validation mode."""
        B, H, Wd, _ = x_bhwc.shape
        if _MODE == 0:
            return ops.conv3x3(x_bhwc, w, _empty((B, H, Wd, w.shape[0]), torch.bfloat16, x_bhwc))
        return ops.conv3x3_f32(x_bhwc, w, _empty((B, H, Wd, w.shape[2]), torch.float32, x_bhwc))

    def _conv_rows(self, x_rows, w, halo):
        """Convolution of a row shard [B, Li, L, C]: `halo(x)` returns the shard with one neighbour row
        above and below ([B, Li+2, L, C], zeros at the image border); rows 1..Li of the result are exact."""
        if halo is None:
            return self._conv(x_rows, w)
        y = self._conv(halo(x_rows), w)
        return y[:, 1:-1].contiguous()

    @torch.no_grad()
    def forward(self, msa, pair, att):
        if pair.shape[1] * pair.shape[2] <= 1:
            # the reference's InstanceNorm2d (:453) refuses a 1 x 1 map with exactly this error
            raise ValueError(f"Expected more than 1 spatial element when training, got input size "
                             f"{torch.Size([pair.shape[0], pair.shape[3], pair.shape[1], pair.shape[2]])}")
        return self._forward_rows(msa, pair, att, 0, pair.shape[1], None, None)

    def _project(self, msa):
        """The first two steps of proj_msa (LN -> Linear, :434-437) on any slice of the MSA: [B,n,l,D] ->
        float32 [B,n,l,d_proj]. Per token, so the long-protein path runs it on residue shards and gathers the
        32-channel result instead of the 384-channel MSA."""
        msa = _as_f32(msa).contiguous()
        pk = self._pack()
        T, D = msa.numel() // msa.shape[-1], msa.shape[-1]
        xn = _ln_into(msa.view(T, D), self.proj_msa[0], _empty((T, D), _bdt(), msa))
        mraw = _empty((T, self.d_proj), torch.float32, msa)
        ops.gemm(xn, pk["Wproj"], cview(mraw), bias=pk["bproj"])
        return mraw.view(*msa.shape[:-1], self.d_proj)

    def _forward_rows(self, msa, pair_rows, att_rows, lo, hi, halo, allreduce, mraw=None):
        """Rows [lo, hi) of the updated pair map. msa: the full MSA (or None with `mraw` = `_project` of the
        full MSA); pair_rows / att_rows: rows [lo, hi) of pair and of the tied attention map. `halo` /
        `allreduce` (long-protein path, sharded.py): neighbour-row exchange for the 3x3 convolutions and the
        sum of the InstanceNorm statistics over the row shards; None for the whole map."""
        pair, att = _as_f32(pair_rows).contiguous(), _as_f32(att_rows).contiguous()
        pk = self._pack()
        if mraw is None:
            mraw = self._project(msa)
        msa = mraw  # device / dtype template for the buffers below
        B, N, L, Q = mraw.shape
        mraw = mraw.reshape(B * N * L, Q)
        Li = hi - lo
        P, H = self.d_pair, self.n_heads
        T, TP = B * N * L, B * Li * L
        adt, bdt = _adt(), _bdt()
        # third step of proj_msa (LN, :438); m kept in float32 (tiny), operand copy for GEMMs
        m32 = _ln_into(mraw, self.proj_msa[2], _empty((T, Q), torch.float32, msa))
        m_op = m32 if _MODE == 1 else _ln_into(mraw, self.proj_msa[2], _empty((T, Q), bdt, msa))
        w = self.poswise_weight._weights(m_op.view(B, N, L, Q))  # [B,N,L,1] (:469-470)
        # outer-product sum operands + msa_1d (:472-482)
        Np = _up8(N)
        xt = _empty((B, L * Q, Np), bdt, msa)
        yt = _empty((B, L * Q, Np), bdt, msa)
        msa1d = _empty((B, L, 2 * Q), torch.float32, msa)
        ops.opm_prep(m32.view(B, N, L, Q), w.view(B, N, L), xt[..., :N], yt[..., :N], msa1d)
        coevol = self.outer_product_mean._run(xt[:, lo * Q:hi * Q, :N], yt[..., :N], B, L)  # f32 [TP, P]
        # feature buffer [coevol_ln | ln_pair | att] — the 716-wide concat is never built (:487-496)
        KF = 2 * P + H
        feat = _empty((TP, _up8(KF)), bdt, msa)
        _ln_into(coevol, self.ln_coevol_feat, feat[:, :P])
        _ln_into(pair.view(TP, P), self.ln_pair, feat[:, P:2 * P])
        ops.convert_rows(att.view(TP, H), feat[:, 2 * P:KF])
        # rank-1 row / column terms of the Linear
        rowt = _empty((B, L, P), torch.float32, msa)
        colt = _empty((B, L, P), torch.float32, msa)
        ops.gemm(msa1d.view(B * L, 2 * Q), pk["Wr"], cview(rowt.view(B * L, P)))
        ops.gemm(msa1d.view(B * L, 2 * Q), pk["Wc"], cview(colt.view(B * L, P)))
        h = _empty((B, Li, L, P), torch.float32, msa)
        ops.gemm(feat.view(B, Li * L, -1)[..., :KF], pk["Wf"][None], h.view(1, 1, B, Li, L, 1, P),
                 bias=pk["bf"],
                 r0=rowt[:, lo:hi].reshape(1, 1, B, Li, 1, 1, P).expand(1, 1, B, Li, L, 1, P),
                 r1=colt.view(1, 1, B, 1, L, 1, P).expand(1, 1, B, Li, L, 1, P))
        # Residual(conv -> IN -> ELU -> conv -> IN) then ELU (:449-462)
        fn = self.resnet[1].fn
        scale = Li / L  # statistics are sums over the whole map: rescale so the kernels' 1/positions applies

        def stats_of(c):
            st = torch.zeros((B, 2, P), dtype=torch.float64, device=msa.device)
            ops.channel_stats(c, st)
            if allreduce is not None:
                allreduce(st)
                st.mul_(scale)
            return st

        h_op = h if _MODE == 1 else ops.convert_rows(h.view(TP, P), _empty((TP, P), adt, msa)).view(B, Li, L, P)
        c1 = self._conv_rows(h_op, pk["conv1"], halo).view(B, Li * L, P)
        a1 = ops.instnorm_apply(c1, stats_of(c1), pk["g1"], pk["b1"], fn[2].eps, _empty(c1.shape, adt, msa), elu=True)
        c2 = self._conv_rows(a1.view(B, Li, L, P), pk["conv2"], halo).view(B, Li * L, P)
        out = ops.instnorm_apply(c2, stats_of(c2), pk["g2"], pk["b2"], fn[6].eps,
                                 _empty(c2.shape, torch.float32, msa), res=h.view(B, Li * L, P), elu=True)
        return out.view(B, Li, L, P)


# ---------------------------------------------------------------------------------------------
# Pair axial attention (:501-547)
# ---------------------------------------------------------------------------------------------
class PairUpdateWithAxialAttentionLayer(nn.Module):
    def __init__(self, d_pair, d_ff, n_heads, p_dropout, performer_kws):
        super().__init__()
        self.row_attn = PerformerSelfAttention(dim=d_pair, heads=n_heads, dropout=p_dropout,
                                               generalized_attention=True, **performer_kws)
        self.col_attn = PerformerSelfAttention(dim=d_pair, heads=n_heads, dropout=p_dropout,
                                               generalized_attention=True, **performer_kws)
        self.ff = FeedForward(d_pair, d_ff, p_dropout)
        self.layer = nn.Sequential(
            Residual(nn.Sequential(nn.LayerNorm(d_pair), RowWise(self.row_attn))),
            Residual(nn.Sequential(nn.LayerNorm(d_pair), ColWise(self.col_attn))),
            Residual(nn.Sequential(nn.LayerNorm(d_pair), self.ff)),
        )

    def _run(self, x4):
        x4 = _performer_block(self.layer[0].fn[0], self.row_attn, x4, token_dim=1)
        x4 = _performer_block(self.layer[1].fn[0], self.col_attn, x4, token_dim=2)
        D = x4.shape[-1]
        return _ff_block(self.layer[2].fn[0], self.ff, x4.view(-1, D), _adt()).view(x4.shape)

    @torch.no_grad()
    def forward(self, x):
        return self._run(_as_f32(x).contiguous())


class PairUpdateWithAxialAttention(nn.Module):
    def __init__(self, d_pair, d_ff, n_heads, p_dropout, n_encoder_layers, performer_kws={}):
        super().__init__()
        self.layers = nn.ModuleList([
            PairUpdateWithAxialAttentionLayer(d_pair, d_ff, n_heads, p_dropout, performer_kws)
            for _ in range(n_encoder_layers)])

    @torch.no_grad()
    def forward(self, x):
        x = _as_f32(x).contiguous()
        for layer in self.layers:
            x = layer._run(x)
        return x


# ---------------------------------------------------------------------------------------------
# pair -> MSA (:550-610)
# ---------------------------------------------------------------------------------------------
class Symmetrization(nn.Module):
    def __init__(self):
        super().__init__()

    @torch.no_grad()
    def forward(self, x):
        """Standalone form (:550-556) on rfk_pair_symmetrize; the fused path symmetrises inside rfk_pair2att_logits."""
        x = _as_f32(x).contiguous()
        if x.dim() != 4 or x.shape[1] != x.shape[2] or x.shape[3] % 4:
            raise ValueError("Symmetrization: expected [B, L, L, C] with C % 4 == 0")
        return ops.pair_symmetrize(x, torch.empty_like(x))


class MsaUpdateWithPairLayer(nn.Module):
    def __init__(self, d_msa, d_pair, n_heads, p_dropout=0.1):
        super().__init__()
        self.n_heads = n_heads
        self.pair2att = nn.Sequential(Symmetrization(), nn.LayerNorm(d_pair), nn.Linear(d_pair, n_heads),
                                      nn.Dropout(p_dropout), nn.Identity(), nn.Softmax(dim=-1))
        self.msa2value = nn.Sequential(nn.LayerNorm(d_msa), nn.Linear(d_msa, d_msa), nn.Identity())
        self.ff = Residual(nn.Sequential(nn.LayerNorm(d_msa), FeedForward(d_msa, d_msa, p_dropout)),
                           p_dropout=p_dropout)
        self.dropout = nn.Dropout(p_dropout)

    def _pack(self):
        def build():
            ln, lin = self.pair2att[1], self.pair2att[2]
            Wf = (lin.weight * ln.weight[None, :]).detach().float().contiguous()
            bf = (lin.weight @ ln.bias + lin.bias).detach().float().contiguous()
            return dict(Wf=Wf, bf=bf, Wv=_w(self.msa2value[1].weight, _bdt()), bv=_f(self.msa2value[1].bias))
        return _packed(self, build)

    def _run(self, msa, att_p):
        """msa: f32 [B,N,L,D]; att_p: [B,H,L,Lp] softmaxed pair-derived attention (operand dtype)."""
        pk = self._pack()
        B, N, L, D = msa.shape
        H = self.n_heads
        dh = D // H
        T = B * N * L
        adt = _bdt()
        Lp = att_p.shape[-1]
        xn = _ln_into(msa.view(T, D), self.msa2value[0], _empty((T, D), adt, msa))
        vt = _empty((B, H, N * dh, Lp), adt, msa)  # b h (n d) j
        ops.gemm(xn.view(B, N * L, D), pk["Wv"][None],
                 vt.view(B, H, N, dh, Lp)[..., :L].permute(0, 2, 4, 1, 3)[None, None], bias=pk["bv"])
        y = _empty((B, N, L, D), torch.float32, msa)

        def as_out(t):  # [B,N,L,D] -> [1, b, h, 1, i, n, d]
            return t.view(B, N, L, H, dh).permute(0, 3, 2, 1, 4).unsqueeze(2)[None]

        ops.gemm(att_p[..., :L], vt[..., :L], as_out(y), r0=as_out(msa))  # msa + updated (:592-595)
        return _ff_block(self.ff.fn[0], self.ff.fn[1], y.view(T, D), _bdt()).view(B, N, L, D)

    @torch.no_grad()
    def forward(self, msa, pair):
        msa = _as_f32(msa).contiguous()
        att_p = _pair2att([self], _as_f32(pair).contiguous())
        return self._run(msa, att_p[:, : self.n_heads])


def _pair2att(layers, pair):
    """softmax_j(Linear(LN(sym(pair)))) for every layer in one pass over `pair`:
    returns [B, len(layers)*H, L, Lp] in the operand dtype."""
    B, L, _, P = pair.shape
    Wf = torch.cat([l._pack()["Wf"] for l in layers], 0)
    bf = torch.cat([l._pack()["bf"] for l in layers], 0)
    Cn = Wf.shape[0]
    eps = layers[0].pair2att[1].eps
    logits = _empty((B, Cn, L, L), torch.float32, pair)
    ops.pair2att_logits(pair, Wf, bf, eps, logits)
    Lp = _up8(L)
    att_p = _empty((B, Cn, L, Lp), _bdt(), pair)
    ops.softmax_rows(logits.view(B * Cn * L, L), att_p.view(B * Cn * L, Lp)[:, :L])
    return att_p


def _pair2att_rows(layers, rows, cols_t, gather_rows):
    """`_pair2att` for a row-sharded pair map: rows [B,Li,L,P] / cols_t [B,L,Li,P] as in ops.pair2att_logits_rows;
    `gather_rows(x [B,C,Li,L]) -> [B,C,L,L]` concatenates every rank's logit rows (softmax_j needs whole rows
    only, but the attention maps are applied to every MSA row shard, so all ranks need all rows)."""
    B, Li, L, P = rows.shape
    Wf = torch.cat([l._pack()["Wf"] for l in layers], 0)
    bf = torch.cat([l._pack()["bf"] for l in layers], 0)
    Cn = Wf.shape[0]
    part = _empty((B, Cn, Li, L), torch.float32, rows)
    ops.pair2att_logits_rows(rows, cols_t, Wf, bf, layers[0].pair2att[1].eps, part)
    logits = gather_rows(part)
    Lp = _up8(L)
    att_p = _empty((B, Cn, L, Lp), _bdt(), rows)
    ops.softmax_rows(logits.view(B * Cn * L, L), att_p.view(B * Cn * L, Lp)[:, :L])
    return att_p


class MsaUpdateWithPair(nn.Module):
    def __init__(self, d_msa, d_pair, n_heads, n_encoder_layers=4, p_dropout=0.1):
        super().__init__()
        self.encoder_layers = nn.ModuleList([MsaUpdateWithPairLayer(d_msa, d_pair, n_heads, p_dropout)
                                             for _ in range(n_encoder_layers)])

    @torch.no_grad()
    def forward(self, msa, pair):
        msa, pair = _as_f32(msa).contiguous(), _as_f32(pair).contiguous()
        return self._run(msa, lambda chunk: _pair2att(chunk, pair))

    def _run(self, msa, att_of):
        """att_of(layers) -> the attention maps [B, len(layers)*H, L, Lp] of those layers (from the whole pair map,
        or from a row shard: rosettafold_pytorch_b200.sharded)."""
        layers = list(self.encoder_layers)
        # chunks of <= 32 output channels per pass over pair (kernel limit)
        H = layers[0].n_heads
        per = max(1, 32 // H)
        for c0 in range(0, len(layers), per):
            chunk = layers[c0:c0 + per]
            att_p = att_of(chunk)
            for i, layer in enumerate(chunk):
                msa = layer._run(msa, att_p[:, i * H:(i + 1) * H])
        return msa


# ---------------------------------------------------------------------------------------------
# MsaUpdateWithPairAndCoord (:865-920): the MSA update of the three-track blocks (:1044, :1123)
# ---------------------------------------------------------------------------------------------
CA_IDX = 1  # reference :15


class MsaUpdateWithPairAndCoord(nn.Module):
    def __init__(self, d_msa, d_state, d_trfm_inner, d_ff, distance_bins=[8, 12, 16, 20], p_dropout=0.1):
        super().__init__()
        self.distance_bins = distance_bins
        self.n_heads = len(self.distance_bins)
        self.scale = (d_state // self.n_heads) ** -0.5
        self.d_inner = d_trfm_inner
        self.ln_msa = nn.LayerNorm(d_msa)
        self.ln_state = nn.LayerNorm(d_state)
        self.to_q = nn.Linear(d_state, d_trfm_inner * self.n_heads)
        self.to_k = nn.Linear(d_state, d_trfm_inner * self.n_heads)
        self.to_v = nn.Linear(d_msa, d_msa)
        self.ln_out = nn.LayerNorm(d_msa)
        self.to_out = Residual(nn.Sequential(nn.LayerNorm(d_msa), FeedForward(d_msa, d_ff, p_dropout)))

    def _pack(self):
        return _packed(self, lambda: dict(
            Wqk=_w(torch.cat([self.to_q.weight, self.to_k.weight], 0), _bdt()),
            bqk=_f(torch.cat([self.to_q.bias, self.to_k.bias], 0)),
            Wv=_w(self.to_v.weight, _bdt()), bv=_f(self.to_v.bias),
            bins=torch.tensor([float(b) for b in self.distance_bins], dtype=torch.float32,
                              device=self.to_v.weight.device)))

    @torch.no_grad()
    def forward(self, xyz, state, msa):
        """xyz: (B, L, 3, 3) backbone atoms, state: (B, L, d_state), msa: (B, N, L, d_msa) -> msa."""
        msa, state = _as_f32(msa).contiguous(), _as_f32(state).contiguous()
        xyz = _as_f32(xyz).contiguous()
        pk = self._pack()
        B, N, L, D = msa.shape
        H, di = self.n_heads, self.d_inner
        dh = D // H
        T = B * N * L
        adt = _bdt()
        Lp = _up8(L)
        # distance-masked attention map from the state track (:892-913)
        sn = _ln_into(state.view(B * L, -1), self.ln_state, _empty((B * L, state.shape[-1]), adt, msa))
        qk = _empty((B * L, 2 * H * di), adt, msa)
        ops.gemm(sn, pk["Wqk"], cview(qk), bias=pk["bqk"])
        qk5 = qk.view(B, L, 2, H, di)
        logits = _empty((B, H, L, Lp), torch.float32, msa)
        ops.gemm(qk5[:, :, 0].permute(0, 2, 1, 3), qk5[:, :, 1].permute(0, 2, 1, 3),
                 logits[..., :L].as_strided((1, B, H, 1, L, 1, L), (0, H * L * Lp, L * Lp, 0, Lp, 0, 1)),
                 alpha=self.scale)
        ops.dist_mask_logits(xyz[:, :, CA_IDX], pk["bins"], logits[..., :L])
        att = _empty((B, H, L, Lp), adt, msa)
        ops.softmax_rows(logits.view(B * H * L, Lp)[:, :L], att.view(B * H * L, Lp)[:, :L])
        # values from the normalised MSA; the residual stream of this module is LN(msa) (:890, :916)
        mn32 = _ln_into(msa.view(T, D), self.ln_msa, _empty((T, D), torch.float32, msa))
        mn = mn32 if _MODE == 1 else ops.convert_rows(mn32, _empty((T, D), adt, msa))
        vt = _empty((B, H, N * dh, Lp), adt, msa)  # b h (n d) j
        ops.gemm(mn.view(B, N * L, D), pk["Wv"][None],
                 vt.view(B, H, N, dh, Lp)[..., :L].permute(0, 2, 4, 1, 3)[None, None], bias=pk["bv"])
        out = _empty((B, N, L, D), torch.float32, msa)
        ops.gemm(att[..., :L], vt[..., :L], out.view(B, N, L, H, dh).permute(0, 3, 2, 1, 4).unsqueeze(2)[None])
        y = _empty((T, D), torch.float32, msa)
        ops.layernorm(out.view(T, D), _f(self.ln_out.weight), _f(self.ln_out.bias), self.ln_out.eps, y, res=mn32)
        return _ff_block(self.to_out.fn[0], self.to_out.fn[1], y, _bdt()).view(B, N, L, D)


# ---------------------------------------------------------------------------------------------
# TwoTrackBlock (:923-968) and the trunk part of ThreeTrackBlock / FinalBlock (:1037-1041,
# :1116-1120), which run the same four calls in the same order.
# ---------------------------------------------------------------------------------------------
class TwoTrackBlock(nn.Module):
    def __init__(self, d_msa, d_pair, n_encoder_layers, p_dropout=0.1):
        super().__init__()
        self.msa_update_using_self_att = MsaUpdateUsingSelfAttention(
            d_msa=d_msa, d_ff=d_msa * 4, n_heads=12, n_encoder_layers=n_encoder_layers, p_dropout=p_dropout)
        self.pair_update_with_msa = PairUpdateWithMsa(d_pair=d_pair, n_heads=12, d_msa=d_msa, d_proj=32)
        self.pair_update_with_axial_attention = PairUpdateWithAxialAttention(
            d_pair=d_pair, d_ff=d_pair * 4, n_heads=8, p_dropout=p_dropout,
            n_encoder_layers=n_encoder_layers, performer_kws={})
        self.msa_update_with_pair = MsaUpdateWithPair(
            d_msa=d_msa, d_pair=d_pair, n_heads=4, n_encoder_layers=n_encoder_layers, p_dropout=p_dropout)

    @torch.no_grad()
    def forward(self, msa, pair):
        msa, att = self.msa_update_using_self_att(msa)
        pair = self.pair_update_with_msa(msa, pair, att)
        pair = self.pair_update_with_axial_attention(pair)
        msa = self.msa_update_with_pair(msa, pair)
        return msa, pair


class TrunkBlocks(nn.Module):
    """The trunk of a RoseTTAFold model: the (msa, pair) -> (msa, pair) part of every block
    (n_two_track_blocks TwoTrackBlocks + the A-D calls of the three-track and final blocks)."""

    def __init__(self, d_msa=384, d_pair=288, n_blocks=13, n_encoder_layers=4, p_dropout=0.1):
        super().__init__()
        self.blocks = nn.ModuleList([TwoTrackBlock(d_msa, d_pair, n_encoder_layers, p_dropout)
                                     for _ in range(n_blocks)])

    @torch.no_grad()
    def forward(self, msa, pair):
        for blk in self.blocks:
            msa, pair = blk(msa, pair)
        return msa, pair


# ---------------------------------------------------------------------------------------------
# weight transfer from a reference instance (state_dict + the plain-list layers it misses)
# ---------------------------------------------------------------------------------------------
def load_reference_weights(mine: nn.Module, ref: nn.Module):
    """Copy every parameter/buffer of a reference module (same class name) into `mine`,
    including the layers the reference keeps in plain Python lists (:602-605)."""
    sd = dict(ref.state_dict())

    def walk(mod, prefix):
        for name in ("encoder_layers", "blocks"):
            held = mod.__dict__.get(name)
            if isinstance(held, list):
                for i, layer in enumerate(held):
                    for k, v in layer.state_dict().items():
                        sd[f"{prefix}{name}.{i}.{k}"] = v
                    walk(layer, f"{prefix}{name}.{i}.")
        for cname, child in mod.named_children():
            walk(child, f"{prefix}{cname}.")

    walk(ref, "")
    missing, unexpected = mine.load_state_dict(sd, strict=False)
    if missing or unexpected:
        raise RuntimeError(f"weight transfer mismatch: missing={missing[:5]} unexpected={unexpected[:5]}")
    return mine

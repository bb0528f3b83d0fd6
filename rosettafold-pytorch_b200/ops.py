"""Tensor-level wrappers over the C ABI (include/rfk.h).

Every function takes CUDA tensors (or strided views of them), validates shapes/strides, and
enqueues exactly one librfk kernel on the current CUDA stream of the tensors' device. Nothing here
computes with PyTorch: torch is used for device memory and streams only. Calling any op without
librfk.so or on a non-CUDA tensor raises.

Call path: public wrapper (validation) -> `_call` (all tensor arguments on ONE device, that device made
current for the launch) -> `torch.ops.rfk.<op>` (torch custom op, `_torch_ops.py`) -> `_CudaBackend.<op>`
(ctypes marshalling) -> `extern "C" rfk_<op>` in librfk.so.
"""
from __future__ import annotations

import contextlib
import ctypes as C

import torch

from . import _lib
from ._lib import (ACT_ELU, ACT_NONE, ACT_RELU, EPI_BLOCKLN32, EPI_STD, RFK_BF16, RFK_F16, RFK_F32,
                   RfkAddr, RfkFavorDesc, RfkGemmDesc)

__all__ = [
    "ACT_NONE", "ACT_RELU", "ACT_ELU", "EPI_STD", "EPI_BLOCKLN32",
    "gemm", "layernorm", "softmax_rows", "tied_att_symmetrize", "poswise_weight", "opm_prep",
    "pair2att_logits", "channel_stats", "instnorm_apply", "favor_attention", "convert_rows", "dist_mask_logits",
    "conv3x3", "pack_conv3x3_weight", "conv3x3_f32", "pack_conv3x3_weight_f32", "msa_embed", "pair_embed",
]


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return RFK_F32
    if t.dtype == torch.bfloat16:
        return RFK_BF16
    if t.dtype == torch.float16:
        return RFK_F16
    raise TypeError(f"rfk ops take float32, bfloat16 or float16 tensors, got {t.dtype}")


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _f32ptr(t, name):
    if t is None:
        return None
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise TypeError(f"{name} must be a contiguous float32 tensor")
    return C.c_void_p(t.data_ptr())


class _CudaBackend:
    """Calls librfk.so. The only backend the package ships."""

    name = "librfk"

    def __init__(self):
        self.lib = _lib.load()

    @staticmethod
    def _stream(t: torch.Tensor):
        if not t.is_cuda:
            raise RuntimeError("rfk ops run on CUDA tensors only (no CPU path exists)")
        return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)

    # -- gemm ---------------------------------------------------------------------------------
    def gemm(self, a, b, c_view, bias, act, alpha, r0, r1, epi, ln_gamma, ln_beta, ln_eps):
        # (one shape / stride tuple per tensor and slice assignments into the ctypes arrays: this runs ~100 times
        # per block on the host)
        d = RfkGemmDesc()
        ash, ast, bsh, bst = a.shape, a.stride(), b.shape, b.stride()
        d.a, d.b = a.data_ptr(), b.data_ptr()
        d.ab_dtype = _dt(a)
        d.act = act
        d.M, d.N, d.K = ash[3], bsh[3], ash[4]
        # Z[0] is the fastest level = tensor dim 2
        d.Z[0:3] = (ash[2], ash[1], ash[0])
        d.a_zs[0:3] = (ast[2] if ash[2] > 1 else 0, ast[1] if ash[1] > 1 else 0, ast[0] if ash[0] > 1 else 0)
        d.b_zs[0:3] = (bst[2] if bsh[2] > 1 else 0, bst[1] if bsh[1] > 1 else 0, bst[0] if bsh[0] > 1 else 0)
        d.lda, d.ldb = ast[3], bst[3]
        d.bias = None if bias is None else bias.data_ptr()
        d.alpha = alpha
        d.epi = epi
        csh = c_view.shape
        d.MR, d.NR = csh[4], csh[6]
        d.c = c_view.data_ptr()
        d.c_dtype = _dt(c_view)

        def fill(addr: RfkAddr, t, sh):
            st = t.stride()
            addr.zs[0:3] = (st[2] if sh[2] > 1 else 0, st[1] if sh[1] > 1 else 0, st[0] if sh[0] > 1 else 0)
            addr.ms[0:2] = (st[4] if sh[4] > 1 else 0, st[3] if sh[3] > 1 else 0)
            addr.ns[0:2] = (st[6] if sh[6] > 1 else 0, st[5] if sh[5] > 1 else 0)

        fill(d.c_addr, c_view, csh)
        if r0 is not None:
            d.r0, d.r0_dtype = r0.data_ptr(), _dt(r0)
            fill(d.r0_addr, r0, r0.shape)
        if r1 is not None:
            d.r1, d.r1_dtype = r1.data_ptr(), _dt(r1)
            fill(d.r1_addr, r1, r1.shape)
        d.ln_eps = ln_eps
        d.ln_gamma = None if ln_gamma is None else ln_gamma.data_ptr()
        d.ln_beta = None if ln_beta is None else ln_beta.data_ptr()
        _lib.check(self.lib.rfk_gemm(C.byref(d), self._stream(a)), "rfk_gemm")

    def layernorm(self, x, gamma, beta, eps, out, res):
        if res is not None:
            _lib.check(self.lib.rfk_layernorm_residual(
                _ptr(x), _dt(x), x.stride(0), _f32ptr(gamma, "gamma"), _f32ptr(beta, "beta"), eps,
                _ptr(res), res.stride(0), _ptr(out), _dt(out), out.stride(0), x.shape[0], x.shape[1],
                self._stream(x)), "rfk_layernorm_residual")
            return
        _lib.check(self.lib.rfk_layernorm(_ptr(x), _dt(x), x.stride(0), _f32ptr(gamma, "gamma"),
                                          _f32ptr(beta, "beta"), eps, _ptr(out), _dt(out),
                                          out.stride(0), x.shape[0], x.shape[1], self._stream(x)),
                   "rfk_layernorm")

    def dist_mask_logits(self, ca, bins, logits):
        B, H, L, ld = logits.shape[0], logits.shape[1], logits.shape[2], logits.stride(2)
        _lib.check(self.lib.rfk_dist_mask_logits(_ptr(ca), ca.stride(1), _f32ptr(bins, "bins"), H, _ptr(logits),
                                                 ld, B, L, self._stream(logits)), "rfk_dist_mask_logits")

    def softmax_rows(self, x, out):
        _lib.check(self.lib.rfk_softmax_rows(_ptr(x), x.stride(0), _ptr(out), _dt(out), out.stride(0),
                                             x.shape[0], x.shape[1], self._stream(x)),
                   "rfk_softmax_rows")

    def tied_att_symmetrize(self, A, att, att16):
        B, H, L, _ = A.shape
        _lib.check(self.lib.rfk_tied_att_symmetrize(
            _ptr(A), _dt(A), A.stride(2), _ptr(att), _ptr(att16),
            0 if att16 is None else att16.stride(2), B, H, L, self._stream(A)),
            "rfk_tied_att_symmetrize")

    def poswise_weight(self, pq, pk, scale, w_out, q, q_scale, qt, H, dh, stats):
        B, N, L, _ = pk.shape
        _lib.check(self.lib.rfk_poswise_weight_stats(
            _ptr(pq), pq.stride(1), _ptr(pk), pk.stride(2), _dt(pk), scale, _ptr(w_out), _ptr(q),
            0 if q is None else q.stride(2), q_scale, _ptr(qt), 0 if qt is None else _dt(qt), _ptr(stats),
            B, N, L, H, dh, self._stream(pk)), "rfk_poswise_weight_stats")

    def opm_prep(self, m, w, xt, yt, msa1d):
        B, N, L, P = m.shape
        _lib.check(self.lib.rfk_opm_prep(_ptr(m), _ptr(w), _ptr(xt), _ptr(yt), _dt(xt), xt.stride(1),
                                         _ptr(msa1d), B, N, L, P, self._stream(m)), "rfk_opm_prep")

    def pair2att_logits(self, pair, Wf, bf, eps, logits):
        B, L, _, D = pair.shape
        Cn = Wf.shape[0]
        _lib.check(self.lib.rfk_pair2att_logits(_ptr(pair), _ptr(Wf), _ptr(bf), eps, _ptr(logits),
                                                logits.stride(2), B, L, D, Cn, self._stream(pair)),
                   "rfk_pair2att_logits")

    def pair2att_logits_rows(self, rows, cols_t, Wf, bf, eps, logits):
        B, Li, L, D = rows.shape
        _lib.check(self.lib.rfk_pair2att_logits_rows(_ptr(rows), _ptr(cols_t), _ptr(Wf), _ptr(bf), eps, _ptr(logits),
                                                     logits.stride(2), B, Li, L, D, Wf.shape[0],
                                                     self._stream(rows)), "rfk_pair2att_logits_rows")

    def channel_stats(self, x, stats):
        B, P, Cn = x.shape
        _lib.check(self.lib.rfk_channel_stats(_ptr(x), _dt(x), _ptr(stats), B, P, Cn, self._stream(x)),
                   "rfk_channel_stats")

    def instnorm_apply(self, x, stats, gamma, beta, eps, res, elu, out):
        B, P, Cn = x.shape
        _lib.check(self.lib.rfk_instnorm_apply(
            _ptr(x), _dt(x), _ptr(stats), _f32ptr(gamma, "gamma"), _f32ptr(beta, "beta"), eps,
            _ptr(res), 0 if res is None else _dt(res), 1 if elu else 0, _ptr(out), _dt(out), B, P, Cn,
            self._stream(x)), "rfk_instnorm_apply")

    def favor_attention(self, q, k, v, out, proj, kind, heads):
        d = RfkFavorDesc()
        d.q, d.k, d.v, d.out = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
        d.proj = proj.data_ptr()
        d.io_dtype = _dt(q)
        d.kind = kind
        d.m_features = proj.shape[0]
        d.heads = heads
        G1, G0, T, _ = q.shape
        d.tokens = T
        d.G[0], d.G[1] = G0, G1
        d.gs[0], d.gs[1] = q.stride(1), q.stride(0)
        d.ts = q.stride(2)
        d.out_gs[0], d.out_gs[1] = out.stride(1), out.stride(0)
        d.out_ts = out.stride(2)
        _lib.check(self.lib.rfk_favor_attention(C.byref(d), self._stream(q)), "rfk_favor_attention")

    def conv3x3(self, x, w_packed, out, dilation=1):
        B, H, L, Cin = x.shape
        _lib.check(self.lib.rfk_conv3x3_nhwc_dil(_ptr(x), _dt(x), _ptr(w_packed), _ptr(out), _dt(out), B, H, L, Cin,
                                                 out.shape[3], int(dilation), self._stream(x)), "rfk_conv3x3_nhwc_dil")

    def conv3x3_f32(self, x, w_packed, out, dilation=1):
        B, H, L, Cin = x.shape
        _lib.check(self.lib.rfk_conv3x3_nhwc_f32_dil(_ptr(x), _ptr(w_packed), _ptr(out), B, H, L, Cin, out.shape[3],
                                                     int(dilation), self._stream(x)), "rfk_conv3x3_nhwc_f32_dil")

    def pair_symmetrize(self, x, out):
        B, L, _, Cn = x.shape
        _lib.check(self.lib.rfk_pair_symmetrize(_ptr(x), _ptr(out), _dt(x), B, L, Cn, self._stream(x)),
                   "rfk_pair_symmetrize")

    def msa_embed(self, tokens, aa_idx, emb, pos_enc, query_enc, out):
        B, N, L = tokens.shape
        _lib.check(self.lib.rfk_msa_embed(_ptr(tokens), _ptr(aa_idx), _ptr(emb), _ptr(pos_enc), _ptr(query_enc), _ptr(out),
                                          B, N, L, emb.shape[1], self._stream(out)), "rfk_msa_embed")

    def pair_embed(self, seq, aa_idx, table_left, table_right, w_sep, bias, pos_enc_half, out):
        B, L = seq.shape
        _lib.check(self.lib.rfk_pair_embed(_ptr(seq), _ptr(aa_idx), _ptr(table_left), _ptr(table_right), _ptr(w_sep),
                                           _ptr(bias), _ptr(pos_enc_half), _ptr(out), B, L, table_left.shape[1],
                                           self._stream(out)), "rfk_pair_embed")

    def convert_rows(self, x, out):
        _lib.check(self.lib.rfk_convert_rows(_ptr(x), _dt(x), x.stride(0), _ptr(out), _dt(out),
                                             out.stride(0), x.shape[0], x.shape[1], self._stream(x)),
                   "rfk_convert_rows")


_backend = None

# Optional per-op device timing (bench.py's roofline leg): CUDA events recorded on the launching
# stream around every call of the selected ops; durations are read after a synchronize.
_timing = None


def start_timing(names):
    """names: iterable of op names ("gemm", "favor_attention", "layernorm", ...)."""
    global _timing
    _timing = {n: [] for n in names}


def stop_timing():
    """Returns {name: dict(ms=total, calls=n, work=total algorithmic work)}; synchronises."""
    global _timing
    rec, _timing = _timing, None
    torch.cuda.synchronize()
    out = {}
    for name, items in (rec or {}).items():
        shapes = {}  # launches of one op with equal algorithmic work = one kernel shape
        for a, b, w in items:
            e = shapes.setdefault(float(w), dict(ms=0.0, calls=0))
            e["ms"] += a.elapsed_time(b)
            e["calls"] += 1
        out[name] = dict(ms=sum(e["ms"] for e in shapes.values()), calls=len(items),
                         work=float(sum(w for _, _, w in items)), shapes=shapes)
    return out


class _Timed:
    def __init__(self, name, work):
        self.on = _timing is not None and name in _timing
        if self.on:
            self.name, self.work = name, work
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)

    def __enter__(self):
        if self.on:
            self.a.record()

    def __exit__(self, *exc):
        if self.on:
            self.b.record()
            _timing[self.name].append((self.a, self.b, self.work))
        return False


def backend():
    global _backend, _ops
    if _backend is None:
        _backend = _CudaBackend()
        from . import _torch_ops

        _ops = _torch_ops.register(_backend)  # torch.ops.rfk.* -> this backend
    return _backend


_ops = {}
_no_guard = contextlib.nullcontext()


def _call(name, *args):
    """Launch op `name`: every tensor argument must live on the same device (a module left on the CPU with CUDA
    inputs would otherwise hand host pointers to a kernel), and that device is made current for the launch (the
    C side configures kernels and reads the SM count of the CURRENT device). With the librfk backend the call
    goes through the torch custom op `torch.ops.rfk.<name>`; a test backend is called directly."""
    dev = None
    for a in args:
        if isinstance(a, torch.Tensor):
            if dev is None:
                dev = a.device
            elif a.device != dev:
                raise RuntimeError(f"rfk.{name}: tensor arguments live on different devices ({dev} and {a.device}); "
                                   "move the module and its inputs to one CUDA device")
    b = backend()
    if b.name != "librfk":
        return getattr(b, name)(*args)
    if dev is None or dev.type != "cuda":
        raise RuntimeError("rfk ops run on CUDA tensors only (no CPU path exists)")
    guard = _no_guard if dev.index == torch.cuda.current_device() else torch.cuda.device(dev)
    with guard:
        return _ops[name](*args)


def _set_backend_for_tests(b):
    """tests/ only: swap in the oracle-backed emulation to check host logic without a GPU."""
    global _backend
    prev = _backend
    _backend = b
    return prev


# ---------------------------------------------------------------------------------------------
# public wrappers (shape / stride validation common to every backend)
# ---------------------------------------------------------------------------------------------
def _lead(t: torch.Tensor, nd: int) -> torch.Tensor:
    """Pad with leading size-1 dims (one view op: this wrapper runs ~100 times per block on the host)."""
    k = nd - t.dim()
    return t.view((1,) * k + tuple(t.shape)) if k > 0 else t


def gemm(a, b, c_view, *, bias=None, act=ACT_NONE, alpha=1.0, r0=None, r1=None, epi=EPI_STD,
         ln_gamma=None, ln_beta=None, ln_eps=1e-5):
    """C[z][m][n] = epi(alpha * sum_k a[z][m][k] b[z][n][k]).

    a: (..Z, M, K) and b: (..Z, N, K) with up to three leading batch dims (b may have size-1 dims
    that broadcast); the last dim must be contiguous. c_view / r0 / r1 are 7-D strided views
    indexed [Z2, Z1, Z0, M1, MR, N1, NR] with m = M1*MR + mr, n = N1*NR + nr (use .expand for
    broadcast addends): the kernel writes/reads exactly those strides.
    """
    if a.dim() > 5 or b.dim() > 5 or c_view.dim() != 7:
        raise ValueError("gemm: a, b must be <=5-D and c_view 7-D")
    a, b = _lead(a, 5), _lead(b, 5)
    if a.dtype != b.dtype:
        raise TypeError("gemm: a and b dtypes differ")
    if a.stride(4) != 1 or b.stride(4) != 1:
        raise ValueError("gemm: K must be the contiguous dimension of a and b")
    ash, bsh, csh = tuple(a.shape), tuple(b.shape), tuple(c_view.shape)
    if ash[4] != bsh[4]:
        raise ValueError(f"gemm: K mismatch {a.shape} vs {b.shape}")
    for i in range(3):
        if bsh[i] != 1 and bsh[i] != ash[i]:
            raise ValueError("gemm: b batch dims must match a or be 1")
    Zs, M, N = ash[:3], ash[3], bsh[3]
    for name, t in (("c_view", c_view), ("r0", r0), ("r1", r1)):
        if t is None:
            continue
        tsh = csh if t is c_view else tuple(t.shape)
        if tsh[:3] != Zs or tsh[3] * tsh[4] != M or tsh[5] * tsh[6] != N:
            raise ValueError(f"gemm: {name} shape {tsh} does not match Z={Zs} M={M} N={N}")
        if tsh[3:] != csh[3:]:
            raise ValueError("gemm: residual views must share c_view's (M1,MR,N1,NR) split")
    if bias is not None and (bias.dtype != torch.float32 or bias.numel() != N or not bias.is_contiguous()):
        raise ValueError("gemm: bias must be contiguous float32 of length N")
    name = "gemm_f32" if a.dtype == torch.float32 else "gemm_bf16"  # (bucket of the 16-bit tensor-core GEMMs)
    with _Timed(name, 2.0 * Zs[0] * Zs[1] * Zs[2] * M * N * ash[4]):  # algorithmic FLOPs
        _call("gemm", a, b, c_view, bias, act, float(alpha), r0, r1, epi, ln_gamma, ln_beta, float(ln_eps))
    return c_view


def cview(t: torch.Tensor) -> torch.Tensor:
    """A 2-D [M, N] tensor (row stride arbitrary) as the trivial 7-D gemm view."""
    M, N = t.shape
    return t.as_strided((1, 1, 1, 1, M, 1, N), (0, 0, 0, 0, t.stride(0), 0, t.stride(1)))


def layernorm(x, gamma, beta, eps, out, res=None):
    """out = LayerNorm(x) (+ res): `res` is an optional float32 addend of the same shape."""
    if x.dim() != 2 or out.dim() != 2 or x.shape != out.shape:
        raise ValueError("layernorm: x and out must be 2-D with equal shapes")
    if x.stride(1) != 1 or out.stride(1) != 1:
        raise ValueError("layernorm: last dim must be contiguous")
    if res is not None and (res.shape != x.shape or res.dtype != torch.float32 or res.stride(1) != 1):
        raise ValueError("layernorm: res must be float32, shaped like x, last dim contiguous")
    nbytes = float(x.numel() * x.element_size() + out.numel() * out.element_size() + (0 if res is None else res.numel() * 4))
    with _Timed("layernorm", nbytes):  # bytes
        _call("layernorm", x, gamma, beta, float(eps), out, res)
    return out


def dist_mask_logits(ca, bins, logits):
    """logits[b,h,i,j] += -1e9 where |ca[b,i] - ca[b,j]| >= bins[h] (reference :899-913).
    ca: float32 [B, L, 3] (any residue stride), bins: float32 [H], logits: float32 [B,H,L,>=L] view."""
    if ca.dim() != 3 or ca.shape[2] != 3 or ca.dtype != torch.float32 or ca.stride(2) != 1 or ca.stride(0) != ca.shape[1] * ca.stride(1):
        raise ValueError("dist_mask_logits: ca must be float32 [B, L, 3] with a uniform residue stride")
    if logits.dim() != 4 or logits.dtype != torch.float32 or logits.stride(3) != 1 or logits.shape[1] != bins.numel():
        raise ValueError("dist_mask_logits: logits must be float32 [B, H, L, L] rows")
    if logits.stride(1) != logits.shape[2] * logits.stride(2) or logits.stride(0) != logits.shape[1] * logits.stride(1):
        raise ValueError("dist_mask_logits: logits must be a row-padded contiguous [B,H,L,ld] buffer")
    _call("dist_mask_logits", ca, bins, logits)
    return logits


def softmax_rows(x, out):
    if x.dim() != 2 or out.shape != x.shape or x.dtype != torch.float32:
        raise ValueError("softmax_rows: x must be 2-D float32 and out the same shape")
    if x.stride(1) != 1 or out.stride(1) != 1:
        raise ValueError("softmax_rows: last dim must be contiguous")
    _call("softmax_rows", x, out)
    return out


def tied_att_symmetrize(A, att, att16=None):
    """A: [B,H,L,L] view (row stride free); att: contiguous f32 [B,L,L,H]; att16: [B,L,L,H] view
    of a wider bf16 buffer (stride(2) = row stride)."""
    B, H, L, L2 = A.shape
    if L != L2 or tuple(att.shape) != (B, L, L, H) or not att.is_contiguous() or att.dtype != torch.float32:
        raise ValueError("tied_att_symmetrize: bad shapes")
    if A.stride(3) != 1 or A.stride(1) != L * A.stride(2) or A.stride(0) != H * A.stride(1):
        raise ValueError("tied_att_symmetrize: A must be [B,H,L,ld] packed")
    _call("tied_att_symmetrize", A, att, att16)
    return att


def poswise_weight(pq, pk, scale, *, w_out=None, q=None, q_scale=1.0, qt=None, heads, d_head, stats=None):
    """pq: [B,L,H*dh] view, pk/q: [B,N,L,H*dh] views (last dim contiguous, rows packed over
    (b,n,l)); w_out: contiguous f32 [B,N,L,H]; qt: contiguous [B,H,L,N*dh]; stats: optional contiguous f32
    [B,L,H,2] receiving (max, sum of exp) of the logits over the N sequences given (sequence-sharded MSAs)."""
    B, N, L, D = pk.shape
    if D != heads * d_head or tuple(pq.shape) != (B, L, D):
        raise ValueError("poswise_weight: bad shapes")
    if pq.dtype != pk.dtype or (q is not None and q.dtype != pk.dtype):
        raise TypeError("poswise_weight: pq, pk, q dtypes differ")
    for t in (pk, q):
        if t is not None and (t.stride(3) != 1 or t.stride(1) != L * t.stride(2) or t.stride(0) != N * t.stride(1)):
            raise ValueError("poswise_weight: pk/q rows must be packed over (b,n,l)")
    if pq.stride(2) != 1 or pq.stride(0) != L * pq.stride(1):
        raise ValueError("poswise_weight: pq rows must be packed over (b,l)")
    if w_out is not None and (tuple(w_out.shape) != (B, N, L, heads) or not w_out.is_contiguous()):
        raise ValueError("poswise_weight: w_out must be contiguous [B,N,L,H]")
    if qt is not None and (tuple(qt.shape) != (B, heads, L, N * d_head) or not qt.is_contiguous()):
        raise ValueError("poswise_weight: qt must be contiguous [B,H,L,N*dh]")
    if stats is not None and (tuple(stats.shape) != (B, L, heads, 2) or not stats.is_contiguous() or
                              stats.dtype != torch.float32):
        raise ValueError("poswise_weight: stats must be contiguous f32 [B,L,H,2]")
    _call("poswise_weight", pq, pk, float(scale), w_out, q, float(q_scale), qt, heads, d_head, stats)


def opm_prep(m, w, xt, yt, msa1d):
    B, N, L, P = m.shape
    if not (m.is_contiguous() and w.is_contiguous() and msa1d.is_contiguous()):
        raise ValueError("opm_prep: m, w, msa1d must be contiguous")
    if m.dtype != torch.float32 or w.dtype != torch.float32 or msa1d.dtype != torch.float32:
        raise TypeError("opm_prep: m, w, msa1d are float32")
    if w.numel() != B * N * L or tuple(msa1d.shape) != (B, L, 2 * P):
        raise ValueError("opm_prep: bad shapes")
    for t in (xt, yt):
        if tuple(t.shape) != (B, L * P, N) or t.stride(2) != 1 or t.stride(0) != L * P * t.stride(1):
            raise ValueError("opm_prep: xt/yt must be [B, L*P, N] views with packed rows")
    if xt.stride(1) != yt.stride(1) or xt.dtype != yt.dtype:
        raise ValueError("opm_prep: xt and yt must share leading dimension and dtype")
    _call("opm_prep", m, w, xt, yt, msa1d)


def pair2att_logits(pair, Wf, bf, eps, logits):
    B, L, L2, D = pair.shape
    Cn = Wf.shape[0]
    if L != L2 or not pair.is_contiguous() or pair.dtype != torch.float32:
        raise ValueError("pair2att_logits: pair must be contiguous f32 [B,L,L,D]")
    if tuple(Wf.shape) != (Cn, D) or tuple(bf.shape) != (Cn,) or not Wf.is_contiguous():
        raise ValueError("pair2att_logits: bad weight shapes")
    if tuple(logits.shape) != (B, Cn, L, L) or logits.dtype != torch.float32 or logits.stride(3) != 1 \
            or logits.stride(1) != L * logits.stride(2) or logits.stride(0) != Cn * logits.stride(1):
        raise ValueError("pair2att_logits: logits must be a packed [B,C,L,ld] f32 view")
    _call("pair2att_logits", pair, Wf, bf, float(eps), logits)
    return logits


def pair2att_logits_rows(rows, cols_t, Wf, bf, eps, logits):
    """Row-sharded form of pair2att_logits: rows f32 [B,Li,L,D] = this rank's rows of the pair map, cols_t f32
    [B,L,Li,D] = the same rows' columns (cols_t[b,j,il] = pair[b,j,i0+il]); logits f32 [B,C,Li,ld] view."""
    B, Li, L, D = rows.shape
    Cn = Wf.shape[0]
    if not rows.is_contiguous() or rows.dtype != torch.float32 or not cols_t.is_contiguous() \
            or cols_t.dtype != torch.float32 or tuple(cols_t.shape) != (B, L, Li, D):
        raise ValueError("pair2att_logits_rows: rows [B,Li,L,D] and cols_t [B,L,Li,D] must be contiguous f32")
    if tuple(Wf.shape) != (Cn, D) or tuple(bf.shape) != (Cn,) or not Wf.is_contiguous():
        raise ValueError("pair2att_logits_rows: bad weight shapes")
    if tuple(logits.shape) != (B, Cn, Li, L) or logits.dtype != torch.float32 or logits.stride(3) != 1 \
            or logits.stride(1) != Li * logits.stride(2) or logits.stride(0) != Cn * logits.stride(1):
        raise ValueError("pair2att_logits_rows: logits must be a packed [B,C,Li,ld] f32 view")
    _call("pair2att_logits_rows", rows, cols_t, Wf, bf, float(eps), logits)
    return logits


def channel_stats(x, stats):
    B, P, Cn = x.shape
    if not x.is_contiguous() or tuple(stats.shape) != (B, 2, Cn) or stats.dtype != torch.float64:
        raise ValueError("channel_stats: x contiguous [B,P,C], stats f64 [B,2,C] (pre-zeroed)")
    _call("channel_stats", x, stats)
    return stats


def instnorm_apply(x, stats, gamma, beta, eps, out, *, res=None, elu=False):
    if not x.is_contiguous() or not out.is_contiguous() or x.shape != out.shape:
        raise ValueError("instnorm_apply: x/out must be contiguous and equal-shaped")
    if res is not None and (res.shape != x.shape or not res.is_contiguous()):
        raise ValueError("instnorm_apply: res must match x")
    if stats.dtype != torch.float64 or not stats.is_contiguous():
        raise ValueError("instnorm_apply: stats must be the contiguous f64 [B,2,C] buffer of channel_stats")
    _call("instnorm_apply", x, stats, gamma, beta, float(eps), res, bool(elu), out)
    return out


def favor_attention(q, k, v, out, proj, *, kind, heads):
    """q,k,v,out: [G1, G0, T, heads*64] strided views (same strides for q,k,v; last dim
    contiguous); proj: contiguous f32 [m, 64]; kind 0 = softmax kernel, 1 = ReLU kernel."""
    if q.dim() != 4 or q.shape != k.shape or q.shape != v.shape or q.shape != out.shape:
        raise ValueError("favor_attention: q,k,v,out must be equal-shaped 4-D views")
    if q.shape[3] != heads * 64:
        raise ValueError("favor_attention: dim_head is fixed at 64")
    if not (q.stride() == k.stride() == v.stride()) or q.stride(3) != 1 or out.stride(3) != 1:
        raise ValueError("favor_attention: q,k,v must share strides, last dim contiguous")
    if not (q.dtype == k.dtype == v.dtype == out.dtype):
        raise TypeError("favor_attention: dtype mismatch")
    if proj.dtype != torch.float32 or not proj.is_contiguous() or proj.shape[1] != 64:
        raise ValueError("favor_attention: proj must be contiguous f32 [m,64]")
    # algorithmic FLOPs: feature maps 2*(2*T*h*64*m) + context/output 2*(2*T*h*64*m)
    tokens_total = q.shape[0] * q.shape[1] * q.shape[2]
    with _Timed("favor_attention", 8.0 * tokens_total * heads * 64 * proj.shape[0]):
        _call("favor_attention", q, k, v, out, proj, int(kind), int(heads))
    return out


def pack_conv3x3_weight(w: torch.Tensor, dtype=torch.bfloat16) -> torch.Tensor:
    """nn.Conv2d weight [Cout, Cin, 3, 3] -> 16-bit [Cout, 9, Cpad] (tap-major, channels padded to a
    multiple of 64 with zeros): the K-major B operand of the implicit GEMM, in the dtype of the image it meets."""
    Cout, Cin = w.shape[:2]
    cpad = (Cin + 63) // 64 * 64
    out = torch.zeros((Cout, 9, cpad), dtype=dtype, device=w.device)
    out[:, :, :Cin] = w.detach().permute(0, 2, 3, 1).reshape(Cout, 9, Cin).to(dtype)
    return out


def conv3x3(x, w_packed, out, dilation=1):
    """3x3 'same' convolution (optionally dilated) without bias on a channels-last map: x bf16 / f16 [B,H,W,Cin]
    contiguous, w_packed from pack_conv3x3_weight in x's dtype, out 16-bit / f32 [B,H,W,Cout] contiguous."""
    if x.dim() != 4 or not x.is_contiguous() or x.dtype not in (torch.bfloat16, torch.float16):
        raise ValueError("conv3x3: x must be contiguous bf16 / f16 [B,H,W,C]")
    Cout, taps, cpad = w_packed.shape
    if taps != 9 or cpad != (x.shape[3] + 63) // 64 * 64 or w_packed.dtype != x.dtype or not w_packed.is_contiguous():
        raise ValueError("conv3x3: w_packed must come from pack_conv3x3_weight for this channel count")
    if tuple(out.shape) != (x.shape[0], x.shape[1], x.shape[2], Cout) or not out.is_contiguous():
        raise ValueError("conv3x3: bad output shape")
    with _Timed("conv3x3", 2.0 * x.shape[0] * x.shape[1] * x.shape[2] * 9 * x.shape[3] * Cout):
        _call("conv3x3", x, w_packed, out, int(dilation))
    return out


def pack_conv3x3_weight_f32(w: torch.Tensor) -> torch.Tensor:
    """nn.Conv2d weight [Cout, Cin, 3, 3] -> f32 [9, Cin, Cout] (tap-major) for the fp32 validation-mode kernel."""
    Cout, Cin = w.shape[:2]
    return w.detach().float().permute(2, 3, 1, 0).reshape(9, Cin, Cout).contiguous()


def conv3x3_f32(x, w_packed, out, dilation=1):
    """fp32 validation-mode 3x3 'same' convolution (dilation 1..8): x f32 [B,H,W,Cin], w_packed from pack_conv3x3_weight_f32,
    out f32 [B,H,W,Cout], all contiguous."""
    if x.dim() != 4 or not x.is_contiguous() or x.dtype != torch.float32:
        raise ValueError("conv3x3_f32: x must be contiguous f32 [B,H,W,C]")
    if w_packed.dim() != 3 or w_packed.shape[0] != 9 or w_packed.shape[1] != x.shape[3] or w_packed.dtype != torch.float32 \
            or not w_packed.is_contiguous():
        raise ValueError("conv3x3_f32: w_packed must come from pack_conv3x3_weight_f32 for this channel count")
    if tuple(out.shape) != (*x.shape[:3], w_packed.shape[2]) or not out.is_contiguous() or out.dtype != torch.float32:
        raise ValueError("conv3x3_f32: bad output")
    _call("conv3x3_f32", x, w_packed, out, int(dilation))
    return out


def pair_symmetrize(x, out):
    """out[b,i,j,:] = 0.5 (x[b,i,j,:] + x[b,j,i,:]) on a contiguous channels-last [B,L,L,C] map (PredictionHead :1166);
    f32 (C % 4 == 0) or 16-bit (C % 8 == 0), out of x's dtype and shape, not aliasing x."""
    if x.dim() != 4 or x.shape[1] != x.shape[2] or not x.is_contiguous():
        raise ValueError("pair_symmetrize: x must be contiguous [B,L,L,C]")
    if out.shape != x.shape or out.dtype != x.dtype or not out.is_contiguous() or out.data_ptr() == x.data_ptr():
        raise ValueError("pair_symmetrize: out must be a distinct contiguous tensor of x's shape and dtype")
    if x.shape[3] % (4 if x.dtype == torch.float32 else 8):
        raise ValueError("pair_symmetrize: channel count must fill 16-byte chunks")
    _call("pair_symmetrize", x, out)
    return out


def _check_index(name, idx, shape, limit):
    """Integer gather indices: contiguous int64 of the given shape with values in [0, limit) (checked like nn.Embedding
    does; the check reads the tensor back, once per forward of an embedding)."""
    if idx.dtype != torch.int64 or tuple(idx.shape) != tuple(shape) or not idx.is_contiguous():
        raise ValueError(f"{name} must be a contiguous int64 tensor of shape {tuple(shape)}")
    if idx.numel() and (int(idx.min()) < 0 or int(idx.max()) >= limit):
        raise IndexError(f"{name}: index out of range [0, {limit})")


def _f32c(name, t, shape):
    if t.dtype != torch.float32 or tuple(t.shape) != tuple(shape) or not t.is_contiguous():
        raise ValueError(f"{name} must be a contiguous float32 tensor of shape {tuple(shape)}")


def msa_embed(tokens, aa_idx, emb, pos_enc, query_enc, out):
    """MsaEmbedding.forward (reference :114-120): out[b,n,l] = emb[tokens[b,n,l]] + pos_enc[aa_idx[b,l]] + query_enc[n>0]."""
    B, N, L = tokens.shape
    V, D = emb.shape
    _check_index("tokens", tokens, (B, N, L), V)
    _check_index("aa_idx", aa_idx, (B, L), pos_enc.shape[0])
    _f32c("pos_enc", pos_enc, (pos_enc.shape[0], D))
    _f32c("query_enc", query_enc, (2, D))
    _f32c("emb", emb, (V, D))
    _f32c("out", out, (B, N, L, D))
    _call("msa_embed", tokens, aa_idx, emb, pos_enc, query_enc, out)
    return out


def pair_embed(seq, aa_idx, table_left, table_right, w_sep, bias, pos_enc_half, out):
    """PairEmbedding.forward without template (reference :147-175) from the per-vocabulary tables of the split Linear
    (include/rfk.h): out[b,i,j] = table_left[seq[b,j]] + table_right[seq[b,i]] + w_sep log(|aa_i - aa_j| + 1) + bias + PE."""
    B, L = seq.shape
    V, D = table_left.shape
    _check_index("seq", seq, (B, L), V)
    _check_index("aa_idx", aa_idx, (B, L), pos_enc_half.shape[0])
    _f32c("table_left", table_left, (V, D))
    _f32c("table_right", table_right, (V, D))
    _f32c("w_sep", w_sep, (D,))
    _f32c("bias", bias, (D,))
    _f32c("pos_enc_half", pos_enc_half, (pos_enc_half.shape[0], D // 2))
    _f32c("out", out, (B, L, L, D))
    _call("pair_embed", seq, aa_idx, table_left, table_right, w_sep, bias, pos_enc_half, out)
    return out


def convert_rows(x, out):
    if x.dim() != 2 or x.shape != out.shape or x.stride(1) != 1 or out.stride(1) != 1:
        raise ValueError("convert_rows: 2-D views with contiguous last dim")
    _call("convert_rows", x, out)
    return out

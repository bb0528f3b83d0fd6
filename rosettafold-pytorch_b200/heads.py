"""Prediction heads of the reference: `PredictionHead`
(rosettafold_pytorch.py:1130-1172) and the dilated 2-D `ResNet` / `ResBlock2D` it is built from (resnet.py:14-83), as
drop-in `nn.Module`s with the reference's constructor signatures and state_dict keys, running on librfk:

  * LayerNorm + Linear projection of the pair map (:1135-1140): rfk_layernorm + rfk_gemm, channels-last throughout (the
    reference's `b i j c -> b c i j` rearranges and their inverses do not exist);
  * symmetrisation 0.5 (p + p^T) for the distance / omega heads (:1166): rfk_pair_symmetrize;
  * 1 x 1 convolutions (resnet.py:61, :80) as GEMMs over positions; InstanceNorm2d + ELU (+ residual) through
    rfk_channel_stats / rfk_instnorm_apply, exactly as in the trunk's convolution block (:451-462);
  * the dilated 3 x 3 convolutions (resnet.py:19-37, dilations 1, 2, 4, 8) as implicit GEMMs on the tcgen05 kernel
    (rfk_conv3x3_nhwc_dil: the nine TMA boxes are shifted by the dilation, zero fill outside the image), SIMT fp32 in the
    validation mode.
Numerics (tensor-core mode): every convolution / GEMM INPUT of a head follows a LayerNorm -> Linear or an
InstanceNorm -> ELU, i.e. it is range-bounded, so the operands are IEEE half (modules._bdt()); the
convolution OUTPUTS in front of an InstanceNorm have whatever scale the weights give them, so they are written in float32
(never rounded to 16 bits), as is the residual stream. Measured against the unmodified reference: 7e-4 ... 8e-4 in this
form, 7e-3 with bf16 operands and bf16 convolution outputs (eight convolutions in series). Forwards are eval-mode (dropout =
identity), no autograd.
"""
import torch
from torch import nn

from . import modules as M
from . import ops
from .ops import cview


class Residual(nn.Module):
    """resnet.py:6-12 (container: the residual add is fused into the InstanceNorm kernel)."""

    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def forward(self, x):
        return self.fn(x) + x


class ResBlock2D(nn.Module):
    """resnet.py:14-44. `forward` takes the reference's [B, C, H, W] layout; the fused path (`_run`) is channels-last."""

    def __init__(self, channel, kernel_size, dilation, p_dropout=0.15):
        super().__init__()
        if kernel_size != 3:
            raise NotImplementedError("ResBlock2D: the reference only builds 3 x 3 blocks (resnet.py:70)")
        self.layer = Residual(
            nn.Sequential(
                nn.Conv2d(channel, channel, kernel_size, dilation=dilation, padding="same", bias=False),
                nn.InstanceNorm2d(channel, affine=True, eps=1e-6),
                nn.ELU(),
                nn.Dropout(p_dropout),
                nn.Conv2d(channel, channel, kernel_size, dilation=dilation, padding="same", bias=False),
                nn.InstanceNorm2d(channel, affine=True, eps=1e-6),
            )
        )
        self.channel, self.dilation = channel, dilation

    def _pack(self):
        def build():
            fn = self.layer.fn
            if M._MODE == 0:
                def pack(w):
                    return ops.pack_conv3x3_weight(w, M._bdt())
            else:
                pack = ops.pack_conv3x3_weight_f32
            return dict(w1=pack(fn[0].weight), w2=pack(fn[4].weight), g1=M._f(fn[1].weight), b1=M._f(fn[1].bias),
                        g2=M._f(fn[5].weight), b2=M._f(fn[5].bias))
        return M._packed(self, build)

    def _run(self, h, h_op):
        """h: float32 [B,H,W,C] residual stream; h_op: its operand-dtype copy. Returns the same pair for the output."""
        pk = self._pack()
        fn = self.layer.fn
        B, H, W, Cn = h.shape
        c1 = _conv(h_op, pk["w1"], self.dilation).view(B, H * W, Cn)
        a1 = ops.instnorm_apply(c1, _stats(c1), pk["g1"], pk["b1"], fn[1].eps, M._empty(c1.shape, M._bdt(), h), elu=True)
        c2 = _conv(a1.view(B, H, W, Cn), pk["w2"], self.dilation).view(B, H * W, Cn)
        out = ops.instnorm_apply(c2, _stats(c2), pk["g2"], pk["b2"], fn[5].eps, M._empty(c2.shape, torch.float32, h),
                                 res=h.view(B, H * W, Cn), elu=True).view(B, H, W, Cn)
        return out, _operand(out)

    @torch.no_grad()
    def forward(self, x):
        h = M._as_f32(x).permute(0, 2, 3, 1).contiguous()
        out, _ = self._run(h, _operand(h))
        return out.permute(0, 3, 1, 2)


def _conv(x_op, w, dilation):
    """Dilated 3x3 convolution of an operand-dtype channels-last map; float32 output in both modes."""
    B, H, W, _ = x_op.shape
    if M._MODE == 0:
        return ops.conv3x3(x_op, w, M._empty((B, H, W, w.shape[0]), torch.float32, x_op), dilation)
    return ops.conv3x3_f32(x_op, w, M._empty((B, H, W, w.shape[2]), torch.float32, x_op), dilation)


def _stats(c):
    st = torch.zeros((c.shape[0], 2, c.shape[2]), dtype=torch.float64, device=c.device)
    ops.channel_stats(c, st)
    return st


def _operand(x32):
    """Operand-dtype copy of a float32 channels-last tensor (the tensor itself in the fp32 mode)."""
    if M._MODE == 1:
        return x32
    Cn = x32.shape[-1]
    return ops.convert_rows(x32.view(-1, Cn), M._empty((x32.numel() // Cn, Cn), M._bdt(), x32)).view(x32.shape)


class ResNet(nn.Module):
    """resnet.py:47-83: 1 x 1 input projection + InstanceNorm + ELU, dilated residual blocks, 1 x 1 output projection."""

    def __init__(self, n_res_blocks, in_channels, intermediate_channels, out_channels, dilations=[1, 2, 4, 8],
                 p_dropout=0.15):
        super().__init__()
        layers = [
            nn.Conv2d(in_channels, intermediate_channels, 1, bias=False),
            nn.InstanceNorm2d(intermediate_channels, affine=True, eps=1e-6),
            nn.ELU(),
        ]
        for block_idx in range(n_res_blocks):
            layers.append(ResBlock2D(intermediate_channels, kernel_size=3,
                                     dilation=dilations[block_idx % len(dilations)], p_dropout=p_dropout))
        layers.append(nn.Conv2d(intermediate_channels, out_channels, 1))
        self.layer = nn.Sequential(*layers)
        self.n_res_blocks = n_res_blocks
        self.in_channels, self.mid_channels, self.out_channels = in_channels, intermediate_channels, out_channels

    def _pack(self):
        def build():
            adt = M._bdt()
            first, last = self.layer[0], self.layer[-1]
            return dict(Win=M._w(first.weight.reshape(self.mid_channels, self.in_channels), adt),
                        g=M._f(self.layer[1].weight), b=M._f(self.layer[1].bias),
                        Wout=M._w(last.weight.reshape(self.out_channels, self.mid_channels), adt), bout=M._f(last.bias))
        return M._packed(self, build)

    def _run(self, x_op):
        """x_op: operand-dtype channels-last [B,H,W,Cin]. Returns float32 [B,H,W,Cout]."""
        pk = self._pack()
        B, H, W, Cin = x_op.shape
        T, Cm = B * H * W, self.mid_channels
        c0 = M._empty((B, H * W, Cm), torch.float32, x_op)
        ops.gemm(x_op.view(T, Cin), pk["Win"], cview(c0.view(T, Cm)))
        h = ops.instnorm_apply(c0, _stats(c0), pk["g"], pk["b"], self.layer[1].eps,
                               M._empty(c0.shape, torch.float32, x_op), elu=True).view(B, H, W, Cm)
        h_op = _operand(h)
        for blk in list(self.layer)[3:-1]:
            h, h_op = blk._run(h, h_op)
        out = M._empty((T, self.out_channels), torch.float32, x_op)
        ops.gemm(h_op.view(T, Cm), pk["Wout"], cview(out), bias=pk["bout"])
        return out.view(B, H, W, self.out_channels)

    @torch.no_grad()
    def forward(self, x):
        """x: [B, Cin, H, W] as in the reference; returns [B, Cout, H, W]."""
        h = M._as_f32(x).permute(0, 2, 3, 1).contiguous()
        if h.shape[1] * h.shape[2] <= 1:
            raise ValueError(f"Expected more than 1 spatial element when training, got input size {torch.Size(x.shape)}")
        return self._run(_operand(h)).permute(0, 3, 1, 2)


class PredictionHead(nn.Module):
    """rosettafold_pytorch.py:1130-1172: pair [B,L,L,C] -> {"theta", "phi", "dist", "omega"} logits [B,L,L,bins]."""

    def __init__(self, in_channels, n_res_blocks, p_dropout):
        super().__init__()
        mid = in_channels
        # (nn.Identity stands where the reference has an einops Rearrange: no parameters, same Sequential indices)
        self.proj = nn.Sequential(nn.LayerNorm(in_channels), nn.Linear(in_channels, mid), nn.Dropout(p_dropout), nn.Identity())
        self.dist_head = nn.Sequential(ResNet(n_res_blocks, in_channels, mid, 37, p_dropout=p_dropout), nn.Identity())
        self.omega_head = nn.Sequential(ResNet(n_res_blocks, in_channels, mid, 37, p_dropout=p_dropout), nn.Identity())
        self.theta_head = nn.Sequential(ResNet(n_res_blocks, in_channels, mid, 37, p_dropout=p_dropout), nn.Identity())
        self.phi_head = nn.Sequential(ResNet(n_res_blocks, in_channels, mid, 19, p_dropout=p_dropout), nn.Identity())
        self.in_channels = in_channels

    def _pack(self):
        def build():
            return dict(W=M._w(self.proj[1].weight, M._bdt()), b=M._f(self.proj[1].bias))
        return M._packed(self.proj, build)

    @torch.no_grad()
    def forward(self, pair):
        pair = M._as_f32(pair).contiguous()
        B, L, L2, Cn = pair.shape
        if L != L2 or Cn != self.in_channels:
            raise ValueError("PredictionHead: pair must be [B, L, L, in_channels]")
        if L * L <= 1:
            raise ValueError(f"Expected more than 1 spatial element when training, got input size "
                             f"{torch.Size([B, Cn, L, L])}")
        pk = self._pack()
        T = B * L * L
        xn = M._ln_into(pair.view(T, Cn), self.proj[0], M._empty((T, Cn), M._bdt(), pair))
        p = M._empty((B, L, L, Cn), torch.float32, pair)
        ops.gemm(xn, pk["W"], cview(p.view(T, Cn)), bias=pk["b"])
        psym = ops.pair_symmetrize(p, torch.empty_like(p))
        p_op, psym_op = _operand(p), _operand(psym)
        return {
            "theta": self.theta_head[0]._run(p_op),
            "phi": self.phi_head[0]._run(p_op),
            "dist": self.dist_head[0]._run(psym_op),
            "omega": self.omega_head[0]._run(psym_op),
        }

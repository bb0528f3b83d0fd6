"""torch custom-op registration of the librfk launchers: `torch.ops.rfk.<op>`.

BASELINE.json's north star names the boundary "a torch custom op over extern "C" launchers": every tensor-level
op of `ops.py` is defined here once in the `rfk` namespace of the PyTorch dispatcher (schema with its mutated
outputs annotated) and implemented for the CUDA dispatch key by the ctypes call into librfk.so (`ops._CudaBackend`,
which marshals pointers / strides / the current stream into the C ABI of include/rfk.h). There is deliberately NO
implementation for any other dispatch key: calling an op on CPU tensors fails inside the dispatcher ("could not run
'rfk::gemm' with arguments from the 'CPU' backend") — the product has no CPU path.

librfk.so itself stays free of torch symbols (plain C ABI); the registration lives on the
Python side (`torch.library`), which is also where the reference's own host code lives.
"""
from __future__ import annotations

import torch

NAMESPACE = "rfk"

# op name -> schema; argument order = the argument order of the matching `_CudaBackend` method
SCHEMAS = {
    "gemm": "(Tensor a, Tensor b, Tensor(a!) c_view, Tensor? bias, int act, float alpha, Tensor? r0, Tensor? r1, "
            "int epi, Tensor? ln_gamma, Tensor? ln_beta, float ln_eps) -> ()",
    "layernorm": "(Tensor x, Tensor? gamma, Tensor? beta, float eps, Tensor(a!) out, Tensor? res) -> ()",
    "dist_mask_logits": "(Tensor ca, Tensor bins, Tensor(a!) logits) -> ()",
    "softmax_rows": "(Tensor x, Tensor(a!) out) -> ()",
    "tied_att_symmetrize": "(Tensor A, Tensor(a!) att, Tensor(b!)? att16) -> ()",
    "poswise_weight": "(Tensor pq, Tensor pk, float scale, Tensor(a!)? w_out, Tensor? q, float q_scale, "
                      "Tensor(b!)? qt, int H, int dh, Tensor(c!)? stats) -> ()",
    "opm_prep": "(Tensor m, Tensor w, Tensor(a!) xt, Tensor(b!) yt, Tensor(c!) msa1d) -> ()",
    "pair2att_logits": "(Tensor pair, Tensor Wf, Tensor bf, float eps, Tensor(a!) logits) -> ()",
    "pair2att_logits_rows": "(Tensor rows, Tensor cols_t, Tensor Wf, Tensor bf, float eps, Tensor(a!) logits) -> ()",
    "channel_stats": "(Tensor x, Tensor(a!) stats) -> ()",
    "instnorm_apply": "(Tensor x, Tensor stats, Tensor gamma, Tensor beta, float eps, Tensor? res, bool elu, "
                      "Tensor(a!) out) -> ()",
    "favor_attention": "(Tensor q, Tensor k, Tensor v, Tensor(a!) out, Tensor proj, int kind, int heads) -> ()",
    "conv3x3": "(Tensor x, Tensor w_packed, Tensor(a!) out, int dilation) -> ()",
    "conv3x3_f32": "(Tensor x, Tensor w_packed, Tensor(a!) out, int dilation) -> ()",
    "pair_symmetrize": "(Tensor x, Tensor(a!) out) -> ()",
    "convert_rows": "(Tensor x, Tensor(a!) out) -> ()",
    "msa_embed": "(Tensor tokens, Tensor aa_idx, Tensor emb, Tensor pos_enc, Tensor query_enc, Tensor(a!) out) -> ()",
    "pair_embed": "(Tensor seq, Tensor aa_idx, Tensor table_left, Tensor table_right, Tensor w_sep, Tensor bias, "
                  "Tensor pos_enc_half, Tensor(a!) out) -> ()",
}

_library = None
_handles = {}


def register(backend):
    """Define the ops (once per process) and bind their CUDA implementation to `backend`'s methods.
    Returns {name: OpOverload}."""
    global _library
    if _library is None:
        lib = torch.library.Library(NAMESPACE, "DEF")
        for name, schema in SCHEMAS.items():
            lib.define(name + schema)
            lib.impl(name, getattr(backend, name), "CUDA")
        _library = lib  # keep alive: the registrations die with the Library object
        ns = getattr(torch.ops, NAMESPACE)
        for name in SCHEMAS:
            _handles[name] = getattr(ns, name).default
    return _handles

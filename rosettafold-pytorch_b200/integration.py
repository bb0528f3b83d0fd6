"""Swap the trunk of a *reference* model for the B200 modules, in place.

Works on any reference block that owns the four trunk children (`TwoTrackBlock`,
`ThreeTrackBlock`, `FinalBlock`, rosettafold_pytorch.py:923-1127) and on a whole `RoseTTAFold`.
Hyper-parameters are read off the reference instance, weights are copied (including the layers the
reference keeps in plain lists), everything else of the model is left untouched.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import modules as M

TRUNK_CHILDREN = ("msa_update_using_self_att", "pair_update_with_msa",
                  "pair_update_with_axial_attention", "msa_update_with_pair")


def _make(name: str, ref: nn.Module) -> nn.Module:
    if name == "msa_update_using_self_att":
        l0 = ref.residue_wise_encoder_layers[0]
        return M.MsaUpdateUsingSelfAttention(
            d_msa=l0.ln.normalized_shape[0], d_ff=l0.ff.fn[1].net[0].out_features,
            n_heads=l0.attn.n_heads, n_encoder_layers=len(ref.residue_wise_encoder_layers))
    if name == "pair_update_with_msa":
        return M.PairUpdateWithMsa(
            d_msa=ref.proj_msa[0].normalized_shape[0], d_proj=ref.proj_msa[1].out_features,
            d_pair=ref.ln_pair.normalized_shape[0],
            n_heads=ref.resnet[0].in_features - 2 * ref.ln_pair.normalized_shape[0] - 4 * ref.proj_msa[1].out_features)
    if name == "pair_update_with_axial_attention":
        l0 = ref.layers[0]
        return M.PairUpdateWithAxialAttention(
            d_pair=l0.ff.net[0].in_features, d_ff=l0.ff.net[0].out_features, n_heads=l0.row_attn.heads,
            p_dropout=0.0, n_encoder_layers=len(ref.layers))
    if name == "msa_update_with_pair":
        l0 = ref.encoder_layers[0]
        return M.MsaUpdateWithPair(
            d_msa=l0.msa2value[1].in_features, d_pair=l0.pair2att[1].normalized_shape[0],
            n_heads=l0.pair2att[2].out_features, n_encoder_layers=len(ref.encoder_layers))
    if name == "msa_update_with_pair_and_coord":  # three-track / final blocks only (:1028-1035)
        return M.MsaUpdateWithPairAndCoord(
            d_msa=ref.ln_msa.normalized_shape[0], d_state=ref.ln_state.normalized_shape[0],
            d_trfm_inner=ref.to_q.out_features // len(ref.distance_bins),
            d_ff=ref.to_out.fn[1].net[0].out_features, distance_bins=list(ref.distance_bins))
    raise KeyError(name)


def _home_device(module: nn.Module):
    for t in list(module.parameters(recurse=True)) + list(module.buffers(recurse=True)):
        return t.device
    return None


def _move_inputs_hook(module: nn.Module, args, kwargs):
    """Forward pre-hook: tensor arguments that live elsewhere are copied to the device of `module`'s weights."""
    dev = _home_device(module)
    if dev is None:
        return None

    def mv(x):
        if isinstance(x, torch.Tensor) and x.device != dev:
            # host -> device may overlap; device -> host must have landed before the CPU code reads it
            return x.to(dev, non_blocking=dev.type == "cuda")
        if isinstance(x, (tuple, list)):
            return type(x)(mv(y) for y in x)
        return x

    return tuple(mv(a) for a in args), {k: mv(v) for k, v in kwargs.items()}


def _add_hop(module: nn.Module) -> None:
    if not getattr(module, "_rfk_hop", False):
        module.register_forward_pre_hook(_move_inputs_hook, with_kwargs=True)
        module._rfk_hop = True


def accelerate_block(block: nn.Module, device=None, hop: bool = False) -> nn.Module:
    """Replace the trunk children of one reference block. `hop`: see `accelerate`."""
    names = TRUNK_CHILDREN + (("msa_update_with_pair_and_coord",) if hasattr(block, "msa_update_with_pair_and_coord") else ())
    for name in names:
        old = getattr(block, name)
        new = M.load_reference_weights(_make(name, old), old).eval()
        setattr(block, name, new.to(device) if device is not None else new)
    if hop:
        # every direct child gets its tensor arguments on its own device: the swapped modules pull msa / pair / xyz
        # onto the GPU, the reference's structure track and heads pull what they consume back to the CPU
        for child in block.children():
            _add_hop(child)
    return block


def _make_embedding(name: str, ref: nn.Module) -> nn.Module:
    from . import embeddings as E

    if name == "msa_emb":  # MsaEmbedding (:106-120)
        return E.MsaEmbedding(d_input=ref.to_embedding.num_embeddings, d_msa=ref.to_embedding.embedding_dim,
                              max_len=ref.pos_enc.max_len)
    use_template = bool(ref.use_template)  # PairEmbedding (:123-181)
    d_pair = ref.proj.out_features
    return E.PairEmbedding(d_input=ref.embed_seq.num_embeddings, d_pair=d_pair, max_len=ref.pos_enc.max_len,
                           use_template=use_template,
                           d_template=ref.proj.in_features - d_pair - 1 if use_template else 64)


def _make_head(ref: nn.Module) -> nn.Module:
    """PredictionHead (:1130-1172) with the reference's widths: in_channels from the LayerNorm, the number of residual
    blocks from the ResNet's Sequential (3 input layers + blocks + 1 output projection, resnet.py:59-81)."""
    from . import heads as H

    resnet = ref.dist_head[0]
    return H.PredictionHead(in_channels=ref.proj[0].normalized_shape[0], n_res_blocks=len(resnet.layer) - 4,
                            p_dropout=ref.proj[2].p)


def _make_graph_block(ref: nn.Module) -> nn.Module:
    """GraphTransformerBlock (:667-677) with the reference block's widths."""
    from . import graph as G

    a = ref.attn
    return G.GraphTransformerBlock(d_node_in=a.node_to_q.in_features, d_node_out=a.node_to_q.out_features // a.n_heads,
                                   d_edge=a.edge_emb.in_features, n_heads=a.n_heads, p_dropout=a.att_dropout.p)


def accelerate(model: nn.Module, device=None, hop: bool = False, embeddings: bool = True, heads: bool = True) -> nn.Module:
    """Replace the trunk of every block of a reference RoseTTAFold (rosettafold_pytorch.py:1220-1267) and, with
    `embeddings` (default), the two embeddings that feed it (`msa_emb`, `pair_emb`, :1205-1219): the reference's
    own embeddings cannot run on a GPU at all (CPU-resident tables gathered by Python loops).

    `hop=True` is for the UNMODIFIED reference model, which only runs on the CPU (its embeddings index CPU tables
    with Python loops and its list-held layers ignore `.to()`): the swapped trunk
    modules live on `device`, everything else stays where it is, and forward pre-hooks copy tensor arguments to the
    device of the module that consumes them. Tensors then cross PCIe only where the trunk meets the reference's own
    code: embeddings -> first block, trunk -> SE(3) structure track (msa, pair), structure track -> coordinate-
    conditioned MSA update (xyz, state), last block -> prediction head. No structure (attribute names,
    `state_dict` keys) changes.

    `heads` (default): `prediction_head` (:1269-1271, :1287) is replaced as well (PredictionHead / ResNet on librfk,
    heads.py; the einops Rearrange layers of the reference hold no state, so the state_dict keys are unchanged), and so
    are the dense GraphTransformerBlocks inside `initial_coord_generation_with_msa_and_pair` (graph.py)."""
    blocks = list(getattr(model, "two_track_blocks", [])) + list(getattr(model, "three_track_blocks", []))
    if hasattr(model, "final_block"):
        blocks.append(model.final_block)
    whole_model = bool(blocks)
    if not blocks and all(hasattr(model, n) for n in TRUNK_CHILDREN):
        blocks = [model]
    for blk in blocks:
        accelerate_block(blk, device, hop)
    if embeddings and whole_model:
        for name in ("msa_emb", "pair_emb"):
            old = getattr(model, name, None)
            if old is not None and not type(old).__module__.startswith("rosettafold_pytorch_b200"):
                new = _make_embedding(name, old)
                new.load_state_dict(old.state_dict(), strict=True)
                setattr(model, name, new.eval().to(device) if device is not None else new.eval())
    if heads and whole_model:
        old = getattr(model, "prediction_head", None)
        if old is not None and not type(old).__module__.startswith("rosettafold_pytorch_b200"):
            new = _make_head(old)
            new.load_state_dict(old.state_dict(), strict=True)
            setattr(model, "prediction_head", new.eval().to(device) if device is not None else new.eval())
    if heads and whole_model:
        # the dense graph-transformer blocks of the initial-coordinate generator (:699-702, :723-724): the reference
        # keeps them in a plain Python list, so they are replaced in that list
        gen = getattr(model, "initial_coord_generation_with_msa_and_pair", None)
        held = gen.__dict__.get("blocks") if gen is not None else None
        if isinstance(held, list):
            for i, old in enumerate(held):
                if type(old).__module__.startswith("rosettafold_pytorch_b200"):
                    continue
                new = _make_graph_block(old)
                new.load_state_dict(old.state_dict(), strict=True)
                held[i] = new.eval().to(device) if device is not None else new.eval()
    if hop and whole_model:
        # the model's own direct children (embeddings, initial coordinates, prediction head); the block containers
        # are skipped, their blocks were handled above
        for child in model.children():
            holds_blocks = isinstance(child, nn.ModuleList) and any(c in blocks for c in child)
            if not holds_blocks and child not in blocks:
                _add_hop(child)
    return model

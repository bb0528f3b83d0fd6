"""ORACLE / TEST INFRASTRUCTURE — deterministic synthetic weights.

Every tensor of a module's state_dict is drawn from a generator seeded by (seed, crc32(key)), so
the same weights are reproduced on any machine from the key names alone (golden fixtures store
outputs and a checksum, not weights). Biases and LayerNorm/InstanceNorm affines are non-trivial
on purpose (default inits would hide bias / affine bugs).
"""
import zlib

import torch

from .performer_ref import gaussian_orthogonal_random_matrix


def synth_state_dict(template: dict, seed: int = 0) -> dict:
    """template: name -> tensor (shapes only are used). Returns name -> float32 CPU tensor."""
    out = {}
    canonical = {}  # aliased entries (PairUpdateWithAxialAttentionLayer registers row_attn/col_attn/ff
    # twice, :505-525) must receive identical values: key them by the first name of the storage
    for name in sorted(template):
        canonical.setdefault((template[name].data_ptr(), tuple(template[name].shape)), name)
    for name in sorted(template):
        shape = tuple(template[name].shape)
        first = canonical[(template[name].data_ptr(), shape)]
        if first != name:
            out[name] = out[first]
            continue
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 31))
        if name.endswith("projection_matrix"):
            t = gaussian_orthogonal_random_matrix(shape[0], shape[1], generator=g)
        elif len(shape) >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = torch.randn(shape, generator=g) * fan_in ** -0.5
        elif name.endswith("weight"):  # LayerNorm / InstanceNorm scale
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:  # biases
            t = 0.1 * torch.randn(shape, generator=g)
        out[name] = t.float()
    return out


def checksum(sd: dict) -> float:
    return float(sum(v.double().abs().sum() for v in sd.values()))


def push_to_reference(ref_module, sd: dict):
    """Load a flat dict (keys as in the b200 modules, i.e. including `encoder_layers.{i}.`) into
    a REFERENCE module, reaching into the plain-list layers its state_dict misses (:602-605)."""
    own = ref_module.state_dict()
    ref_module.load_state_dict({k: sd[k] for k in own}, strict=True)
    used = set(own)

    def walk(mod, prefix):
        for name in ("encoder_layers", "blocks"):
            held = mod.__dict__.get(name)
            if isinstance(held, list):
                for i, layer in enumerate(held):
                    p = f"{prefix}{name}.{i}."
                    lsd = {k: sd[p + k] for k in layer.state_dict()}
                    layer.load_state_dict(lsd, strict=True)
                    used.update(p + k for k in lsd)
                    walk(layer, p)
        for cname, child in mod.named_children():
            walk(child, f"{prefix}{cname}.")

    walk(ref_module, "")
    missing = set(sd) - used
    if missing:
        raise RuntimeError(f"unused synthetic weights: {sorted(missing)[:5]}")
    return ref_module


def synth_inputs(B, N, L, d_msa, d_pair, seed=1234):
    g = torch.Generator().manual_seed(seed)
    msa = torch.randn((B, N, L, d_msa), generator=g)
    pair = torch.randn((B, L, L, d_pair), generator=g)
    return msa, pair

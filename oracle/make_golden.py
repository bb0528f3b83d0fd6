"""ORACLE / TEST INFRASTRUCTURE — generates tests/golden/*.pt from the UNMODIFIED reference
(/root/reference, imported with oracle/shims). Run in the build container only:

    python -m oracle.make_golden

Each fixture holds the outputs of the reference `TwoTrackBlock` stages (A: MSA self-attention,
B: MSA->pair, C: pair axial attention, D: pair->MSA) for synthetic weights/inputs that are
re-derivable from seeds (oracle/weights.py), plus a weight checksum.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import reference_loader as rl  # noqa: E402
from oracle.weights import checksum, push_to_reference, synth_inputs, synth_state_dict  # noqa: E402

CONFIGS = {
    # ragged: N, L not multiples of 8; two layers so "last layer's att" and shared-pair logic matter
    "two_track_small": dict(d_msa=96, d_pair=72, n_layers=2, B=2, N=5, L=20, seed=3),
    # default feature widths (d_msa 384 / d_pair 288), one layer
    "two_track_default": dict(d_msa=384, d_pair=288, n_layers=1, B=1, N=6, L=24, seed=4),
}


COORD_CONFIGS = {
    # MsaUpdateWithPairAndCoord (:865-920) as built by the three-track blocks (:1028-1035), ragged N / L
    "msa_pair_coord": dict(d_msa=96, d_state=32, d_inner=32, d_ff=192, B=2, N=5, L=20, seed=6),
}


def synth_coords(B, L, d_state, seed):
    """A random-walk backbone with ~3.8 A C-alpha spacing (so all four distance bins are exercised)
    and a random state track."""
    g = torch.Generator().manual_seed(seed)
    steps = torch.randn((B, L, 3), generator=g)
    ca = torch.cumsum(3.8 * steps / steps.norm(dim=-1, keepdim=True), dim=1)
    xyz = torch.stack([ca + 1.46 * torch.randn((B, L, 3), generator=g) / 3 ** 0.5, ca,
                       ca + 1.52 * torch.randn((B, L, 3), generator=g) / 3 ** 0.5], dim=2)
    return xyz, torch.randn((B, L, d_state), generator=g)


def make_coord(ref, rf):
    for name, c in COORD_CONFIGS.items():
        mine = rf.MsaUpdateWithPairAndCoord(c["d_msa"], c["d_state"], c["d_inner"], c["d_ff"])
        sd = synth_state_dict(mine.state_dict(), seed=c["seed"])
        rmod = ref.MsaUpdateWithPairAndCoord(c["d_msa"], c["d_state"], c["d_inner"], c["d_ff"])
        rmod.load_state_dict(sd, strict=True)
        rmod.eval()
        msa, _ = synth_inputs(c["B"], c["N"], c["L"], c["d_msa"], 8, seed=c["seed"] + 100)
        xyz, state = synth_coords(c["B"], c["L"], c["d_state"], c["seed"] + 200)
        with torch.no_grad():
            out = rmod(xyz, state, msa)
        path = os.path.join(ROOT, "tests", "golden", f"{name}.pt")
        torch.save(dict(config=c, weight_checksum=checksum(sd), xyz=xyz, state=state, msa_out=out,
                        generator="oracle/make_golden.py on the unmodified reference (CPU fp32, eval)"), path)
        print(name, tuple(out.shape), os.path.getsize(path))


def main():
    import rosettafold_pytorch_b200 as rf

    ref = rl.load()
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    make_coord(ref, rf)
    if "--coord-only" in sys.argv:
        return
    for name, c in CONFIGS.items():
        mine = rf.TwoTrackBlock(c["d_msa"], c["d_pair"], n_encoder_layers=c["n_layers"])
        sd = synth_state_dict(mine.state_dict(), seed=c["seed"])
        torch.manual_seed(0)
        rblk = ref.TwoTrackBlock(c["d_msa"], c["d_pair"], n_encoder_layers=c["n_layers"])
        push_to_reference(rblk, sd)
        rl.fix_eval(rblk)
        msa, pair = synth_inputs(c["B"], c["N"], c["L"], c["d_msa"], c["d_pair"], seed=c["seed"] + 100)
        with torch.no_grad():
            m, att = rblk.msa_update_using_self_att(msa)
            p1 = rblk.pair_update_with_msa(m, pair, att)
            p2 = rblk.pair_update_with_axial_attention(p1)
            m2 = rblk.msa_update_with_pair(m, p2)
            m_full, p_full = rblk(msa, pair)
        assert torch.equal(m_full, m2) and torch.equal(p_full, p2)
        out = dict(config=c, weight_checksum=checksum(sd), msa_a=m, att=att, pair_b=p1, pair_c=p2, msa_d=m2,
                   generator="oracle/make_golden.py on the unmodified reference (CPU fp32, eval)")
        path = os.path.join(ROOT, "tests", "golden", f"{name}.pt")
        torch.save(out, path)
        print(name, {k: tuple(v.shape) for k, v in out.items() if torch.is_tensor(v)}, os.path.getsize(path))


if __name__ == "__main__":
    main()

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


def main():
    import rosettafold_pytorch_b200 as rf

    ref = rl.load()
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
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

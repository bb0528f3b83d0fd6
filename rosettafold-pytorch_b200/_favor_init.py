"""Construction-time initialisation of the FAVOR+ projection buffer (host side, runs once).

Gaussian-orthogonal random features (Choromanski et al., "Rethinking Attention with Performers"):
blocks of orthonormal rows from the QR of a Gaussian matrix, stacked to m rows and rescaled by
the norms of independent Gaussian vectors, so each row is distributed like a d-dim Gaussian while
rows inside a block stay exactly orthogonal. Same construction (and torch RNG call order) as the
`projection_matrix` buffer of performer_pytorch's FastAttention, which is what a reference
checkpoint carries under `...fast_attention.projection_matrix`.
"""
import torch


def gaussian_orthogonal_random_matrix(nb_rows: int, nb_cols: int) -> torch.Tensor:
    n_full = nb_rows // nb_cols
    blocks = []
    for _ in range(n_full):
        q, _ = torch.linalg.qr(torch.randn(nb_cols, nb_cols), mode="reduced")
        blocks.append(q.t())
    rest = nb_rows - n_full * nb_cols
    if rest > 0:
        q, _ = torch.linalg.qr(torch.randn(nb_cols, nb_cols), mode="reduced")
        blocks.append(q.t()[:rest])
    scale = torch.randn(nb_rows, nb_cols).norm(dim=1)
    return scale[:, None] * torch.cat(blocks)

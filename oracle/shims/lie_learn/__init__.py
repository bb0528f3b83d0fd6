"""ORACLE shim for `lie_learn` (an un-vendored dependency of the reference's SE(3) track, absent here).
Only `representations.SO3.wigner_d.wigner_D_matrix` for l <= 2 exists: the closed forms are pinned INSIDE the
reference by its own identities (equivariant_attention/from_se3cnn/SO3.py:153-154, :186-193), which assert
at import of the basis code."""

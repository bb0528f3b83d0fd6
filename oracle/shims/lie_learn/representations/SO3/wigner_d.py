"""ORACLE shim: real Wigner-D matrices for l <= 2 in the basis the reference expects.

  D^1(a,b,c) = A rot(a,b,c) A^T                        (SO3.py:153-154: irr_repr(1) @ A == A @ rot)
  D^2(a,b,c) = T5 (R x R) T5^T (T5 T5^T)^-1            (SO3.py:186-193: irr_repr(2) @ to5 == to5 @ kron(R, R))
with rot = rot_z(a) rot_y(b) rot_z(c) (SO3.py:26-55). The reference needs no higher order (num_degrees = 2)."""
import numpy as np

_A = np.array([[0, 1, 0], [0, 0, 1], [1, 0, 0]], dtype=np.float64)
_T5 = np.array([[0, 1, 0, 1, 0, 0, 0, 0, 0],
                [0, 0, 0, 0, 0, 1, 0, 1, 0],
                [-3 ** .5 / 3, 0, 0, 0, -3 ** .5 / 3, 0, 0, 0, 12 ** .5 / 3],
                [0, 0, 1, 0, 0, 0, 1, 0, 0],
                [1, 0, 0, 0, -1, 0, 0, 0, 0]], dtype=np.float64)
_T5_PINV = _T5.T @ np.linalg.inv(_T5 @ _T5.T)


def _rot_z(g):
    c, s = np.cos(g), np.sin(g)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float64)


def _rot_y(b):
    c, s = np.cos(b), np.sin(b)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float64)


def wigner_D_matrix(l, alpha, beta, gamma, **_):
    r = _rot_z(float(alpha)) @ _rot_y(float(beta)) @ _rot_z(float(gamma))
    l = int(l)
    if l == 0:
        return np.ones((1, 1))
    if l == 1:
        return _A @ r @ _A.T
    if l == 2:
        return _T5 @ np.kron(r, r) @ _T5_PINV
    raise NotImplementedError("lie_learn shim: wigner_D_matrix only up to l = 2")

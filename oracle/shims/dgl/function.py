"""ORACLE shim: the three dgl.function built-ins the reference uses."""


class _Builtin:
    def __init__(self, kind, a, b, out):
        self.kind, self.a, self.b, self.out = kind, a, b, out


def mean(msg, out):
    return _Builtin("mean", msg, None, out)


def sum(msg, out):  # noqa: A001
    return _Builtin("sum", msg, None, out)


def e_dot_v(lhs, rhs, out):
    return _Builtin("e_dot_v", lhs, rhs, out)

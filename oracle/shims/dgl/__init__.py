"""ORACLE shim: import-only stand-in for `dgl` (SE(3) track, out of the hot path). The trunk
oracle never calls into it; anything that does gets a clear error."""
__version__ = "1.1.0"


def graph(*a, **k):
    raise NotImplementedError("dgl is not available: the SE(3) structure track is outside the trunk oracle")

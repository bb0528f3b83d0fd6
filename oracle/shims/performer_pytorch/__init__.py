"""ORACLE shim: `performer_pytorch` is not installable here; see oracle/performer_ref.py
(PARITY UNPINNED)."""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _root not in sys.path:
    sys.path.insert(0, _root)
from oracle.performer_ref import FastAttention, SelfAttention  # noqa: E402,F401

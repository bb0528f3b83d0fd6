"""Importable alias of the package directory `rosettafold-pytorch_b200/` (a hyphen cannot be
imported): `import rosettafold_pytorch_b200` executes that directory's __init__ with its
sub-modules (`ops`, `modules`, `_lib`, ...) resolved from there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "rosettafold-pytorch_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))

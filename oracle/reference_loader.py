"""ORACLE / TEST INFRASTRUCTURE — imports the UNMODIFIED reference from /root/reference with the
dependency shims of oracle/shims on sys.path. Only usable in the build container (the GPU box
has no /root/reference); nothing under tests -m gpu, smoke() or bench.py may call this."""
import os
import sys

REFERENCE_ROOT = os.environ.get("RFK_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "rosettafold_pytorch"))


def load():
    """Returns the reference module `rosettafold_pytorch.rosettafold_pytorch`."""
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    # appended (not prepended): the reference tree has its own top-level `tests/` directory that
    # must not shadow this repository's `tests` package; real installs of the shimmed
    # dependencies, if any, also win over the shims this way
    for p in (REFERENCE_ROOT, _SHIMS):
        if p not in sys.path:
            sys.path.append(p)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import rosettafold_pytorch.rosettafold_pytorch as ref
    return ref


def fix_eval(module):
    """model.eval() misses the plain-list sub-modules (:602-605, :699-702): walk them too."""
    module.eval()
    for sub in list(module.modules()):
        for name in ("encoder_layers", "blocks"):
            held = getattr(sub, name, None)
            if isinstance(held, list):
                for layer in held:
                    fix_eval(layer)
    return module

"""Import the UNMODIFIED reference (idiap/DepthInSpace) from /root/reference.

TEST INFRASTRUCTURE ONLY, and only usable in the build container: /root/reference does
not exist on the GPU box, so nothing marked ``gpu``, smoke() or bench.py may import this
module.  It is used by tests/test_oracle_pinning.py (skipped when the checkout is absent)
and by oracle/gen_golden.py to produce the committed fixtures under tests/golden/.

The reference needs four accommodations to import on a box without its un-vendored
native dependency (Connecting-the-Dots torchext), matplotlib or a GPU:
  * empty ``ext_cpu`` / ``ext_cuda`` modules        (model/ext_functions.py:38-39)
  * a stub ``matplotlib``                            (co/__init__.py:29)
  * ``torch.cuda.synchronize`` as a no-op            (model/networks.py:67,70)
  * ``photometric_loss`` bound to the in-tree ``photometric_loss_pytorch``
    (model/ext_functions.py:156-183; model/networks.py:372 resolves it at call time)
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DIS_REFERENCE_ROOT", "/root/reference")
_CACHE = {}


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "ext_functions.py"))


def load():
    """-> namespace(networks, ext_functions, multi_frame_networks) of the reference."""
    if "ns" in _CACHE:
        return _CACHE["ns"]
    if not available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    import torch

    for name in ("ext_cpu", "ext_cuda"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        m = types.ModuleType("matplotlib")
        m.use = lambda *a, **k: None
        sys.modules["matplotlib"] = m
        sys.modules["matplotlib.pyplot"] = types.ModuleType("matplotlib.pyplot")
    if not torch.cuda.is_available():
        torch.cuda.synchronize = lambda *a, **k: None
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from model import ext_functions, multi_frame_networks, networks

    ext_functions.photometric_loss = ext_functions.photometric_loss_pytorch
    ns = types.SimpleNamespace(networks=networks, ext_functions=ext_functions,
                               multi_frame_networks=multi_frame_networks)
    _CACHE["ns"] = ns
    return ns

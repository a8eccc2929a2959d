"""`ext_cuda` stand-in.  Point config.json's CTD_DIR at .../depthinspace_b200 so that the
reference's `sys.path.append(CTD_DIR/torchext); import ext_cuda` (model/ext_functions.py:35-39)
picks this module up: the reference then runs its only live ext call
(photometric_loss, model/networks.py:372) on libdis_b200.so, unmodified.
"""
import os
import sys

_pkg_parent = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _pkg_parent not in sys.path:
    sys.path.insert(0, _pkg_parent)

from depthinspace_b200.ext_functions import photometric_loss_backward, photometric_loss_forward  # noqa: E402,F401


def _dead(name):
    def f(*a, **k):
        raise NotImplementedError(
            f"ext_cuda.{name}: wrapped by the reference (model/ext_functions.py:41-110) but never called; "
            "not part of the DepthInSpace hot path")
    return f


nn_cuda = _dead("nn_cuda")
crosscheck_cuda = _dead("crosscheck_cuda")
proj_nn_cuda = _dead("proj_nn_cuda")
xcorrvol_cuda = _dead("xcorrvol_cuda")

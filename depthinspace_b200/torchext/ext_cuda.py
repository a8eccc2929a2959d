"""`ext_cuda` stand-in.  Point config.json's CTD_DIR at .../depthinspace_b200 so that the
reference's `sys.path.append(CTD_DIR/torchext); import ext_cuda` (model/ext_functions.py:35-39)
picks this module up: the reference then runs its only live ext call
(photometric_loss, model/networks.py:372) on libdis_b200.so, unmodified.  The four ops the reference wraps but
never calls (nn / crosscheck / proj_nn / xcorrvol, :41-110) are exported too (parity unpinned: their definition is
in the un-vendored CTD torchext).
"""
import os
import sys

_pkg_parent = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _pkg_parent not in sys.path:
    sys.path.insert(0, _pkg_parent)

from depthinspace_b200.ext_functions import (crosscheck_cuda, nn_cuda, photometric_loss_backward,  # noqa: E402,F401
                                             photometric_loss_forward, proj_nn_cuda, xcorrvol_cuda)

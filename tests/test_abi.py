"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
include/dis_b200.h declares; the Python surface validates arguments like the reference does.
No compute calls here (no GPU in this container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "dis_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"DIS_API\s+[\w\s\*]+?\b(dis_\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib_path():
    from depthinspace_b200 import build
    return build.build()


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert len(syms) >= 18
    for must in ("dis_photometric_loss_forward", "dis_photometric_loss_backward", "dis_lcn_forward",
                 "dis_pattern_loss_forward", "dis_smooth_loss_forward", "dis_flow_warp_forward"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"libdis_b200.so lacks {missing}"


def test_ctypes_signatures_cover_header(lib_path):
    from depthinspace_b200 import _lib
    assert set(declared_symbols()) <= set(_lib.SIGNATURES), set(declared_symbols()) - set(_lib.SIGNATURES)
    lib = _lib.load()
    assert lib.dis_abi_version() >= 1
    assert lib.dis_status_string(0) == b"ok"
    assert b"loss type" in lib.dis_status_string(-1)


def test_argument_validation_without_gpu(lib_path):
    """Error paths return before any CUDA call, so they can be exercised on a CPU box."""
    from depthinspace_b200 import _lib
    lib = _lib.load()
    one = ctypes.c_void_p(16)  # never dereferenced on these paths
    assert lib.dis_photometric_loss_forward(one, one, one, 1, 1, 8, 8, 9, 7, 0.5, None) == -1   # invalid type
    assert lib.dis_photometric_loss_forward(one, one, one, 1, 1, 8, 8, 4, 3, 0.5, None) == -3   # even block
    assert lib.dis_photometric_loss_forward(one, one, one, 1, 1, 8, 8, 17, 3, 0.5, None) == -3  # too large
    assert lib.dis_photometric_loss_forward(None, one, one, 1, 1, 8, 8, 9, 3, 0.5, None) == -4
    assert lib.dis_photometric_loss_forward(one, one, one, 1, 0, 8, 8, 9, 3, 0.5, None) == -2
    assert lib.dis_photometric_loss_forward(one, one, one, 0, 1, 8, 8, 9, 3, 0.5, None) == 0    # empty batch
    assert lib.dis_lcn_forward(one, one, one, 1, 8, 8, 9, 0.05, None) == -2                      # radius >= size
    assert lib.dis_sobel_forward(one, one, 1, 8, 8, 4, None) == -6
    # FuseNet gather: tl in 1..8, tidx in range, flows required for tl > 1
    assert lib.dis_flow_warp_gather_forward(one, None, one, 4, 1, 2, 3, 8, 8, None) == -4
    assert lib.dis_flow_warp_gather_forward(one, None, one, 9, 1, 2, 3, 8, 8, None) == -4
    arr = (ctypes.c_void_p * 8)(*[16] * 8)
    assert lib.dis_flow_warp_gather_forward(one, arr, one, 9, 1, 2, 3, 8, 8, None) == -2         # tl > 8
    assert lib.dis_flow_warp_gather_forward(one, arr, one, 4, 4, 2, 3, 8, 8, None) == -2         # tidx out of range
    assert lib.dis_flow_warp_gather_backward(arr, one, one, 4, 0, 0, 3, 8, 8, None) == 0         # empty batch
    assert lib.dis_pattern_loss_num_partials(2, 512, 432) == 2 * 16 * 7
    with pytest.raises(Exception, match="invalid loss type"):
        _lib.check(-1)
    with pytest.raises(_lib.DisB200Error):
        _lib.check(-3)


def test_python_surface_matches_reference_errors():
    import torch
    from depthinspace_b200 import ext_functions, _ops
    assert [_ops.loss_type_id(t) for t in ("mse", "SAD", "census_mse", "Census_SAD")] == [0, 1, 2, 3]
    with pytest.raises(Exception, match="invalid loss type"):
        ext_functions.photometric_loss(torch.zeros(1, 1, 4, 4), torch.zeros(1, 1, 4, 4), 3, type="ssim")
    with pytest.raises(RuntimeError, match="no CPU path"):  # product path never falls back to the CPU
        ext_functions.photometric_loss(torch.zeros(1, 1, 4, 4), torch.zeros(1, 1, 4, 4), 3, type="mse")


def test_ext_cuda_dropin_is_importable():
    """The reference does sys.path.append(CTD_DIR/'torchext'); import ext_cpu; import ext_cuda."""
    import importlib
    import sys
    p = os.path.join(ROOT, "depthinspace_b200", "torchext")
    sys.path.insert(0, p)
    try:
        ext_cpu = importlib.import_module("ext_cpu")
        ext_cuda = importlib.import_module("ext_cuda")
    finally:
        sys.path.remove(p)
    for name in ("photometric_loss_forward", "photometric_loss_backward", "nn_cuda", "crosscheck_cuda",
                 "proj_nn_cuda", "xcorrvol_cuda"):
        assert callable(getattr(ext_cuda, name))
    with pytest.raises(RuntimeError, match="CUDA-only"):
        ext_cpu.photometric_loss_forward()

"""ctypes binding of libdis_b200.so (C-ABI declared in include/dis_b200.h).

There is no CPU fallback: if the shared library has not been built (python -m
depthinspace_b200.build, or __graft_entry__.build()) importing a symbol raises, loudly.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DIS_B200_LIB") or os.path.join(_HERE, "lib", "libdis_b200.so")  # env override: A/B builds

_c = ctypes
_f = _c.c_void_p          # device pointer (float* / int32_t*), nullable
_i = _c.c_int
_fl = _c.c_float
_sz = _c.c_size_t
_st = _c.c_void_p         # cudaStream_t

# name -> argtypes ; every entry point returns int (dis_status) unless listed in _RESTYPE
SIGNATURES = {
    "dis_abi_version": [],
    "dis_status_string": [_i],
    "dis_last_cuda_error": [],
    "dis_lcn_forward": [_f, _f, _f, _i, _i, _i, _i, _fl, _st],
    "dis_lcn_prepare_input": [_f, _f, _f, _i, _i, _i, _i, _i, _fl, _st],
    "dis_lcn_backward": [_f, _f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _fl, _st],
    "dis_photometric_loss_forward": [_f, _f, _f, _i, _i, _i, _i, _i, _i, _fl, _st],
    "dis_photometric_loss_backward": [_f, _f, _f, _f, _i, _i, _i, _i, _i, _i, _fl, _st],
    "dis_pattern_warp_forward": [_f, _f, _f, _f, _f, _f, _i, _i, _i, _st],
    "dis_pattern_loss_num_partials": [_i, _i, _i],
    "dis_pattern_loss_forward": [_f, _f, _f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _fl, _st],
    "dis_pattern_loss_forward_scaled": [_f, _f, _f, _f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _fl, _st],
    "dis_reduce_pairs": [_f, _i, _f, _st],
    "dis_reduce_pairs_batched": [_f, _i, _i, _f, _st],
    "dis_pattern_loss_multi_num_partials": [_i, _i, _i],
    "dis_pattern_loss_point_num_partials": [_i, _i, _i],
    "dis_pattern_loss_point_forward": [_c.POINTER(_c.c_void_p), _i, _f, _f, _f, _c.POINTER(_c.c_void_p), _c.POINTER(_c.c_void_p), _f, _f, _i, _f, _i, _i, _i, _i, _i, _st],
    "dis_pattern_loss_multi_forward": [_c.POINTER(_c.c_void_p), _i, _f, _f, _f, _c.POINTER(_c.c_void_p), _f, _i, _i, _i, _i, _i, _fl, _st],
    "dis_pattern_loss_multi_forward_scaled": [_c.POINTER(_c.c_void_p), _i, _f, _f, _f, _c.POINTER(_c.c_void_p), _f, _f, _i, _i, _i, _i, _i, _fl, _st],
    "dis_scale_by_device_scalar": [_f, _f, _sz, _f, _f, _st],
    "dis_masked_l1_forward": [_f, _f, _f, _fl, _f, _f, _sz, _st],
    "dis_mul": [_f, _f, _f, _sz, _st],
    "dis_l1_num_partials": [_sz],
    "dis_l1_forward": [_f, _f, _f, _f, _sz, _st],
    "dis_sobel_forward": [_f, _f, _i, _i, _i, _i, _st],
    "dis_sobel_backward": [_f, _f, _i, _i, _i, _i, _st],
    "dis_smooth_loss_num_partials": [_i, _i, _i],
    "dis_smooth_loss_forward": [_f, _f, _f, _f, _i, _i, _i, _st],
    "dis_smooth_loss_forward_scaled": [_f, _f, _f, _f, _i, _i, _i, _fl, _i, _st],
    "dis_flow_warp_forward": [_f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _st],
    "dis_flow_warp_backward": [_f, _f, _f, _f, _f, _i, _i, _i, _i, _st],
    "dis_flow_warp_gather_forward": [_f, _c.POINTER(_c.c_void_p), _f, _i, _i, _i, _i, _i, _i, _st],
    "dis_flow_warp_gather_backward": [_c.POINTER(_c.c_void_p), _f, _f, _i, _i, _i, _i, _i, _i, _st],
    "dis_flow_warp_gather_all_forward": [_f, _c.POINTER(_c.c_void_p), _f, _i, _i, _i, _i, _i, _st],
    "dis_flow_warp_gather_all_backward": [_c.POINTER(_c.c_void_p), _f, _f, _i, _i, _i, _i, _i, _st],
    "dis_flow_consistency_num_partials": [_i, _i, _i],
    "dis_flow_consistency_forward": [_f] * 10 + [_i, _f, _f, _f, _fl, _fl, _f, _f, _f, _f, _f, _i, _i, _i, _st],
    "dis_geometric_grad_combine": [_c.POINTER(_c.c_void_p), _c.POINTER(_c.c_int), _i, _f, _f, _fl, _f, _i, _i, _i, _i, _st],
    "dis_ext_nn": [_f, _f, _f, _c.c_int64, _c.c_int64, _i, _st],
    "dis_ext_crosscheck": [_f, _f, _f, _c.c_int64, _c.c_int64, _st],
    "dis_ext_proj_nn": [_f, _f, _f, _f, _i, _i, _i, _i, _st],
    "dis_ext_xcorrvol": [_f, _f, _f, _i, _i, _i, _i, _i, _st],
    "dis_resize_bilinear_forward": [_c.POINTER(_c.c_void_p), _c.POINTER(_c.c_void_p), _i, _i, _i, _i, _i, _i, _i, _i, _st],
    "dis_resize_bilinear_backward": [_f, _f, _i, _i, _i, _i, _i, _i, _st],
    "dis_conv3d_out_size": [_i, _i, _i],
    "dis_conv3d_scratch_elems": [_i, _i, _i, _i],
    "dis_conv3d_gather_forward": [_f, _f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _i, _i, _i, _st],
    "dis_conv3d_gather_backward": [_f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _i, _i, _i, _st],
    "dis_conv3d_rank": [_f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _i, _i, _st],
    "dis_conv3d_gather_features": [_f, _f, _f, _i, _i, _i, _i, _i, _i, _i, _i, _st],
    "dis_combine2": [_f, _f, _f, _sz, _f, _f, _f, _fl, _st],
}
_RESTYPE = {"dis_status_string": _c.c_char_p, "dis_last_cuda_error": _c.c_char_p, "dis_conv3d_scratch_elems": _c.c_size_t}
OPTIONAL = set()

_lib = None


class DisB200Error(Exception):
    pass


def load():
    """Load (once) and return the ctypes library handle."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the sm_100a CUDA library has not been built and there is no CPU "
            "fallback.  Run `python -m depthinspace_b200.build` (needs nvcc).")
    lib = _c.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            if name in OPTIONAL:
                continue
            raise ImportError(f"{LIB_PATH} does not export {name}; rebuild it")
        fn.argtypes = argtypes
        fn.restype = _RESTYPE.get(name, _i)
    _lib = lib
    return lib


LAUNCHES = 0  # kernels of libdis_b200 enqueued through check() by this process (bench.py reports it)


def check(status, launches=1):
    """Raise on a non-zero dis_status (mirrors the reference raising from the ext boundary)."""
    global LAUNCHES
    if status == 0:
        LAUNCHES += launches
        return
    lib = load()
    msg = lib.dis_status_string(int(status)).decode()
    if status == -1:
        raise Exception("invalid loss type")  # model/ext_functions.py:153
    if status == -5:
        msg += ": " + lib.dis_last_cuda_error().decode()
    raise DisB200Error(f"libdis_b200: {msg} (status {status})")

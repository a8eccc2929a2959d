"""`ext_cpu` stand-in: importable (model/ext_functions.py:38 imports it unconditionally) but every
entry point raises -- the B200 implementation deliberately has no CPU path."""


def _no_cpu(name):
    def f(*a, **k):
        raise RuntimeError(f"ext_cpu.{name}: depthinspace_b200 is CUDA-only (sm_100a); move the tensors to the GPU")
    return f


photometric_loss_forward = _no_cpu("photometric_loss_forward")
photometric_loss_backward = _no_cpu("photometric_loss_backward")
nn_cpu = _no_cpu("nn_cpu")
crosscheck_cpu = _no_cpu("crosscheck_cpu")
proj_nn_cpu = _no_cpu("proj_nn_cpu")
xcorrvol_cpu = _no_cpu("xcorrvol_cpu")

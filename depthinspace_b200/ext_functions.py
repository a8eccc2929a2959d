"""Drop-in for the reference's operator boundary, model/ext_functions.py.

Same names, argument meaning and error behaviour as the reference wrappers over the
Connecting-the-Dots `ext_cuda` module, with the CUDA side provided by libdis_b200.so:

    photometric_loss(es, ta, block_size, type='mse', eps=0.1)       (reference :142-154)
    PhotometricLossFunction.forward / .backward                      (reference :115-140)

    nn / crosscheck / proj_nn / xcorrvol                             (reference :41-110; wrapped there, never called:
                                                                      semantics from the un-vendored CTD torchext, parity unpinned)

The reference dispatches CPU tensors to `ext_cpu`; this implementation is CUDA-only and
raises for CPU tensors (no CPU fallback by design).
"""
import torch

from . import _ops


class PhotometricLossFunction(torch.autograd.Function):
    """es, ta [N,C,H,W] -> [N,1,H,W]; gradient w.r.t. es only (reference :140)."""

    @staticmethod
    def forward(ctx, es, ta, block_size, type, eps):
        ctx.save_for_backward(es, ta)
        ctx.block_size = block_size
        ctx.type = type
        ctx.eps = eps
        return _ops.photometric_loss_forward(es, ta, block_size, type, eps)

    @staticmethod
    def backward(ctx, grad_out):
        es, ta = ctx.saved_tensors
        grad_es = _ops.photometric_loss_backward(es, ta, grad_out.contiguous(), ctx.block_size, ctx.type, ctx.eps)
        return grad_es, None, None, None, None


def photometric_loss(es, ta, block_size, type='mse', eps=0.1):
    type_id = _ops.loss_type_id(type)  # raises Exception('invalid loss type') like the reference
    return PhotometricLossFunction.apply(es, ta, block_size, type_id, eps)


# ext_cuda-style free functions (what `import ext_cuda` exposes to the reference, :124, :137)
def photometric_loss_forward(es, ta, block_size, type, eps):
    return _ops.photometric_loss_forward(es, ta, block_size, type, eps)


def photometric_loss_backward(es, ta, grad_out, block_size, type, eps):
    return _ops.photometric_loss_backward(es, ta, grad_out, block_size, type, eps)


# ---- the four wrappers the reference defines but never calls (:41-110); forward only, no gradients, as there ----
def nn(in0, in1):
    return _ops.ext_nn(in0.detach(), in1.detach())


def crosscheck(in0, in1):
    return _ops.ext_crosscheck(in0, in1)


def proj_nn(xyz0, xyz1, K, patch_size):
    return _ops.ext_proj_nn(xyz0.detach(), xyz1.detach(), K.detach(), patch_size)


def xcorrvol(in0, in1, n_disps, block_size):
    return _ops.ext_xcorrvol(in0.detach(), in1.detach(), n_disps, block_size)


nn_cuda, crosscheck_cuda, proj_nn_cuda, xcorrvol_cuda = nn, crosscheck, proj_nn, xcorrvol   # ext_cuda names (:46, 64, 81, 100)

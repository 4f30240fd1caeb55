"""``D3D`` -- the shim a reference maintainer drops in place of the compiled pybind11 extension of 3D/dcn (vision.cpp:4-7).

Same two functions, same positional signatures and return values as the reference's module:

    deform_conv_forward(input, weight, bias, offset, kd, kh, kw, sd, sh, sw, pd, ph, pw, dd, dh, dw,
                        group, deformable_group, im2col_step) -> output                      (deform_conv.h:10-47)
    deform_conv_backward(input, weight, bias, offset, grad_output, kd, ..., im2col_step)
                        -> [grad_input, grad_offset, grad_weight, grad_bias]                 (deform_conv.h:49-92)

so that ``3D/dcn/functions/deform_conv_func.py`` (and its copy under synapse/) runs unchanged: put this directory on
``sys.path`` ahead of the compiled extension (or copy the file next to deform_conv_func.py).  It binds libdlka_b200.so through
ctypes only (no other module of this package is imported), which is exactly the stub INTEGRATION.md section 1 describes;
tests/test_integration_shim.py imports it under the name ``D3D`` and runs the reference's calling sequence through it.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_L = ctypes.CDLL(os.environ.get("DLKA_LIB") or os.path.join(os.path.dirname(_HERE), "libdlka_b200.so"))
_L.dlka_deform_conv3d_workspace_bytes.restype = ctypes.c_size_t
_L.dlka_deform_conv3d_backward_workspace_bytes.restype = ctypes.c_size_t
_L.dlka_status_string.restype = ctypes.c_char_p
_L.dlka_status_string.argtypes = [ctypes.c_int]
_MATH_BF16X3 = 1


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _check(st):
    if st != 0:
        raise RuntimeError(_L.dlka_status_string(st).decode())


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def deform_conv_forward(input, weight, bias, offset, kd, kh, kw, sd, sh, sw, pd, ph, pw, dd, dh, dw,
                        group, deformable_group, im2col_step):
    if not input.is_cuda:
        raise RuntimeError("Not implemented on the CPU")                 # deform_conv.h:46
    if not input.is_contiguous():
        raise RuntimeError("input tensor has to be contiguous")          # deform_conv_cuda.cu:41
    if not weight.is_contiguous():
        raise RuntimeError("weight tensor has to be contiguous")         # deform_conv_cuda.cu:42
    B, C, D, H, W = input.shape
    Co = weight.shape[0]
    Do = (D + 2 * pd - (dd * (kd - 1) + 1)) // sd + 1                     # deform_conv_cuda.cu:78-80
    Ho = (H + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1
    Wo = (W + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1
    out = torch.empty(B, Co, Do, Ho, Wo, device=input.device, dtype=torch.float32)
    n = _L.dlka_deform_conv3d_workspace_bytes(B, C, D, H, W, Co, kd, kh, kw, sd, sh, sw, pd, ph, pw, dd, dh, dw,
                                              group, deformable_group)
    ws = torch.empty(max(n, 1), dtype=torch.uint8, device=input.device)
    with torch.cuda.device(input.device):
        _check(_L.dlka_deform_conv3d_forward(_p(input), _p(weight), _p(bias.contiguous()), _p(offset.contiguous()), _p(out),
                                             B, C, D, H, W, Co, kd, kh, kw, sd, sh, sw, pd, ph, pw, dd, dh, dw,
                                             group, deformable_group, im2col_step, _MATH_BF16X3, _p(ws), ctypes.c_size_t(n),
                                             _stream(input)))
    return out


def deform_conv_backward(input, weight, bias, offset, grad_output, kd, kh, kw, sd, sh, sw, pd, ph, pw, dd, dh, dw,
                         group, deformable_group, im2col_step):
    if not input.is_cuda:
        raise RuntimeError("Not implemented on the CPU")                 # deform_conv.h:84
    B, C, D, H, W = input.shape
    Co = weight.shape[0]
    offset = offset.contiguous()
    gi, go = torch.empty_like(input), torch.empty_like(offset)
    gw, gb = torch.empty_like(weight), torch.empty_like(bias)
    n = _L.dlka_deform_conv3d_backward_workspace_bytes(B, C, D, H, W, Co, kd, kh, kw, sd, sh, sw, pd, ph, pw, dd, dh, dw,
                                                       group, deformable_group)
    ws = torch.empty(max(n, 1), dtype=torch.uint8, device=input.device)
    with torch.cuda.device(input.device):
        _check(_L.dlka_deform_conv3d_backward(_p(input), _p(weight), _p(offset), _p(grad_output.contiguous()),
                                              _p(gi), _p(go), _p(gw), _p(gb), B, C, D, H, W, Co, kd, kh, kw, sd, sh, sw,
                                              pd, ph, pw, dd, dh, dw, group, deformable_group, im2col_step, _MATH_BF16X3,
                                              _p(ws), ctypes.c_size_t(n), _stream(input)))
    return [gi, go, gw, gb]                                               # the reference's order, deform_conv_cuda.cu:282-284

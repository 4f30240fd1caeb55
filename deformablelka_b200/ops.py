"""Functional host API over the C ABI (include/dlka.h).  Tensors are fp32 CUDA tensors; every
function runs on the caller's current CUDA stream and allocates only its output (+ a cached
workspace)."""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import Block2dParams, Block3dParams, DwGeom3d, Workspace, check, dptr, lib, stream_ptr


def _triple(v):
    return tuple(int(i) for i in v) if isinstance(v, (tuple, list)) else (int(v),) * 3


def _pair(v):
    return tuple(int(i) for i in v) if isinstance(v, (tuple, list)) else (int(v),) * 2


def _out_extent(n, pad, dil, k, stride):
    return (n + 2 * pad - (dil * (k - 1) + 1)) // stride + 1


def _math(math) -> int:
    return _lib.default_math() if math is None else _lib.math_mode(math)


# ----------------------------------------------------------------------------------------------
# 3D operator: D3D.deform_conv_forward (3D/dcn/src/deform_conv.h:10-47)
# ----------------------------------------------------------------------------------------------
def deform_conv3d_forward(input: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, offset: torch.Tensor,
                          kernel_size, stride, padding, dilation, group: int, deformable_group: int,
                          im2col_step: int = 64, math=None) -> torch.Tensor:
    """Same argument meaning and error behaviour as ``D3D.deform_conv_forward``."""
    if not input.is_cuda:
        raise RuntimeError("Not implemented on the CPU")  # deform_conv.h:46
    if not input.is_contiguous():
        raise RuntimeError("input tensor has to be contiguous")  # deform_conv_cuda.cu:41
    if not weight.is_contiguous():
        raise RuntimeError("weight tensor has to be contiguous")  # :42
    if input.dim() != 5 or weight.dim() != 5:
        raise RuntimeError("input and weight must be 5-D (NCDHW / OIDHW)")
    kd, kh, kw = _triple(kernel_size)
    sd, sh, sw = _triple(stride)
    pd, ph, pw = _triple(padding)
    dd, dh, dw = _triple(dilation)
    B, C, D, H, W = input.shape
    Co, Cg, kd_, kh_, kw_ = weight.shape
    if (kd_, kh_, kw_) != (kd, kh, kw):
        raise RuntimeError(f"Input shape and kernel shape wont match: ({kh} x {kw} x {kd} vs {kh_} x {kw_} x {kd_}).")  # :73
    if C != Cg * group:
        raise RuntimeError(f"Input shape and kernel channels wont match: ({C} vs {Cg * group}).")  # :76
    if C % group or Co % group:
        raise RuntimeError(f"channels({C}) and channels_out({Co}) must divide group({group})")  # :65
    step = min(B, im2col_step)
    if B % step:
        raise RuntimeError(f"batch({B}) must divide im2col_step({step})")  # :63
    Do, Ho, Wo = _out_extent(D, pd, dd, kd, sd), _out_extent(H, ph, dh, kh, sh), _out_extent(W, pw, dw, kw, sw)
    K = kd * kh * kw
    if tuple(offset.shape) != (B, deformable_group * 3 * K, Do, Ho, Wo):
        raise RuntimeError(f"offset shape {tuple(offset.shape)} does not match "
                           f"{(B, deformable_group * 3 * K, Do, Ho, Wo)}")
    if bias is None or bias.numel() != Co:
        raise RuntimeError("bias must be a tensor with channels_out elements")  # modules/deform_conv.py:39
    offset = offset.contiguous()  # the reference reads raw pointers; callers already pass contiguous
    out = torch.empty(B, Co, Do, Ho, Wo, dtype=torch.float32, device=input.device)
    ws_bytes = lib.dlka_deform_conv3d_workspace_bytes(B, C, D, H, W, Co, kd, kh, kw, sd, sh, sw, pd, ph, pw,
                                                      dd, dh, dw, group, deformable_group)
    ws = Workspace.get(input.device, ws_bytes)
    with torch.cuda.device(input.device):
        st = lib.dlka_deform_conv3d_forward(
            dptr(input, "input"), dptr(weight, "weight"), dptr(bias.contiguous(), "bias"), dptr(offset, "offset"),
            dptr(out), B, C, D, H, W, Co, kd, kh, kw, sd, sh, sw, pd, ph, pw, dd, dh, dw, group, deformable_group,
            im2col_step, _math(math), ws.data_ptr(), ws.numel(), stream_ptr(input.device))
    check(st, "dlka_deform_conv3d_forward")
    return out


def deform_conv3d_sample_indices(offset: torch.Tensor, in_size, kernel_size, stride=1, padding=0, dilation=1,
                                 deformable_group: int = 1) -> Tuple[torch.Tensor, torch.Tensor]:
    D, H, W = in_size
    kd, kh, kw = _triple(kernel_size); sd, sh, sw = _triple(stride); pd, ph, pw = _triple(padding); dd, dh, dw = _triple(dilation)
    offset = offset.contiguous()
    B = offset.shape[0]
    Vo = offset.shape[2] * offset.shape[3] * offset.shape[4]
    K = kd * kh * kw
    low = torch.empty(B * deformable_group, Vo, K, 3, dtype=torch.int32, device=offset.device)
    mask = torch.empty(B * deformable_group, Vo, K, dtype=torch.int32, device=offset.device)
    with torch.cuda.device(offset.device):
        st = lib.dlka_deform_conv3d_sample_indices(dptr(offset, "offset"), low.data_ptr(), mask.data_ptr(), B, D, H, W,
                                                   kd, kh, kw, sd, sh, sw, pd, ph, pw, dd, dh, dw, deformable_group,
                                                   stream_ptr(offset.device))
    check(st, "dlka_deform_conv3d_sample_indices")
    return low, mask


# ----------------------------------------------------------------------------------------------
# 2D operator: torchvision.ops.deform_conv2d (torchvision/ops/deform_conv.py:14-107)
# ----------------------------------------------------------------------------------------------
def deform_conv2d(input: torch.Tensor, offset: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
                  stride=(1, 1), padding=(0, 0), dilation=(1, 1), mask: Optional[torch.Tensor] = None, math=None) -> torch.Tensor:
    """Same signature and semantics as ``torchvision.ops.deform_conv2d``."""
    if not input.is_cuda:
        raise RuntimeError("Not implemented on the CPU (deformablelka_b200 is CUDA-only)")
    sh, sw = _pair(stride); ph, pw = _pair(padding); dh, dw = _pair(dilation)
    B, C, H, W = input.shape
    Co, Cg, kh, kw = weight.shape
    n_off = offset.shape[1] // (2 * kh * kw)
    n_wg = C // Cg
    if n_off == 0:
        raise RuntimeError(
            "the shape of the offset tensor at dimension 1 is not valid. It should "
            "be a multiple of 2 * weight.size[2] * weight.size[3].\n"
            f"Got offset.shape[1]={offset.shape[1]}, while 2 * weight.size[2] * weight.size[3]={2 * kh * kw}")
    Ho, Wo = _out_extent(H, ph, dh, kh, sh), _out_extent(W, pw, dw, kw, sw)
    if tuple(offset.shape) != (B, n_off * 2 * kh * kw, Ho, Wo):
        raise RuntimeError(f"offset shape {tuple(offset.shape)} does not match {(B, n_off * 2 * kh * kw, Ho, Wo)}")
    if mask is not None and tuple(mask.shape) != (B, n_off * kh * kw, Ho, Wo):
        raise RuntimeError(f"mask shape {tuple(mask.shape)} does not match {(B, n_off * kh * kw, Ho, Wo)}")
    input = input.contiguous(); offset = offset.contiguous(); weight = weight.contiguous()
    out = torch.empty(B, Co, Ho, Wo, dtype=torch.float32, device=input.device)
    ws_bytes = lib.dlka_deform_conv2d_workspace_bytes(B, C, H, W, Co, kh, kw, sh, sw, ph, pw, dh, dw, n_wg, n_off)
    ws = Workspace.get(input.device, ws_bytes)
    with torch.cuda.device(input.device):
        st = lib.dlka_deform_conv2d_forward(
            dptr(input, "input"), dptr(weight, "weight"), dptr(offset, "offset"),
            dptr(None if mask is None else mask.contiguous(), "mask"),
            dptr(None if bias is None else bias.contiguous(), "bias"), dptr(out),
            B, C, H, W, Co, kh, kw, sh, sw, ph, pw, dh, dw, n_wg, n_off, _math(math),
            ws.data_ptr(), ws.numel(), stream_ptr(input.device))
    check(st, "dlka_deform_conv2d_forward")
    return out


def deform_conv2d_backward(input, offset, weight, mask, grad_output, stride=(1, 1), padding=(0, 0), dilation=(1, 1),
                           need_bias_grad: bool = False):
    """Gradients of ``torchvision.ops.deform_conv2d`` (row N2, 2D half): returns (grad_input, grad_offset, grad_weight,
    grad_mask or None, grad_bias or None) -- what autograd through torch.ops.torchvision.deform_conv2d returns."""
    if not input.is_cuda:
        raise RuntimeError("Not implemented on the CPU (deformablelka_b200 is CUDA-only)")
    sh, sw = _pair(stride); ph, pw = _pair(padding); dh, dw = _pair(dilation)
    B, C, H, W = input.shape
    Co, Cg, kh, kw = weight.shape
    n_off = offset.shape[1] // (2 * kh * kw)
    n_wg = C // Cg
    Ho, Wo = _out_extent(H, ph, dh, kh, sh), _out_extent(W, pw, dw, kw, sw)
    if tuple(grad_output.shape) != (B, Co, Ho, Wo):
        raise RuntimeError(f"grad_output shape {tuple(grad_output.shape)} does not match {(B, Co, Ho, Wo)}")
    input = input.contiguous(); offset = offset.contiguous(); weight = weight.contiguous(); grad_output = grad_output.contiguous()
    mask = None if mask is None else mask.contiguous()
    gi, go, gw = torch.empty_like(input), torch.empty_like(offset), torch.empty_like(weight)
    gm = None if mask is None else torch.empty_like(mask)
    gb = torch.empty(Co, dtype=torch.float32, device=input.device) if need_bias_grad else None
    ws = Workspace.get(input.device, lib.dlka_deform_conv2d_backward_workspace_bytes(
        B, C, H, W, Co, kh, kw, sh, sw, ph, pw, dh, dw, n_wg, n_off, 0 if mask is None else 1))
    with torch.cuda.device(input.device):
        st = lib.dlka_deform_conv2d_backward(
            dptr(input, "input"), dptr(weight, "weight"), dptr(offset, "offset"), dptr(mask, "mask"), dptr(grad_output, "grad_output"),
            dptr(gi), dptr(gw), dptr(go), dptr(gm), dptr(gb), B, C, H, W, Co, kh, kw, sh, sw, ph, pw, dh, dw, n_wg, n_off,
            ws.data_ptr(), ws.numel(), stream_ptr(input.device))
    check(st, "dlka_deform_conv2d_backward")
    return gi, go, gw, gm, gb


def deform_conv2d_sample_indices(offset, in_size, kernel_size, stride=1, padding=0, dilation=1, n_offset_grps=1):
    H, W = in_size
    kh, kw = _pair(kernel_size); sh, sw = _pair(stride); ph, pw = _pair(padding); dh, dw = _pair(dilation)
    offset = offset.contiguous()
    B = offset.shape[0]
    P = offset.shape[2] * offset.shape[3]
    K = kh * kw
    low = torch.empty(B * n_offset_grps, P, K, 2, dtype=torch.int32, device=offset.device)
    mask = torch.empty(B * n_offset_grps, P, K, dtype=torch.int32, device=offset.device)
    with torch.cuda.device(offset.device):
        st = lib.dlka_deform_conv2d_sample_indices(dptr(offset, "offset"), low.data_ptr(), mask.data_ptr(), B, H, W,
                                                   kh, kw, sh, sw, ph, pw, dh, dw, n_offset_grps, stream_ptr(offset.device))
    check(st, "dlka_deform_conv2d_sample_indices")
    return low, mask


def deform_conv_pack3d(input, offset_weight, offset_bias, weight, bias, stride, padding, dilation, groups,
                       deformable_groups, im2col_step=64, math=None) -> torch.Tensor:
    """DeformConvPack.forward (synapse/deform_conv.py:93-105): conv_offset + deformable conv in one call."""
    if not input.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    if input.dim() != 5:
        raise RuntimeError("expected a 5-D NCDHW tensor")
    input = input.contiguous()
    sd, sh, sw = _triple(stride); pd, ph, pw = _triple(padding); dd, dh, dw = _triple(dilation)
    B, C, D, H, W = input.shape
    Co, Cg, kd, kh, kw = weight.shape
    if C != Cg * groups:
        raise RuntimeError(f"Input shape and kernel channels wont match: ({C} vs {Cg * groups}).")
    step = min(B, im2col_step)
    if B % step:
        raise RuntimeError(f"batch({B}) must divide im2col_step({step})")
    Do, Ho, Wo = _out_extent(D, pd, dd, kd, sd), _out_extent(H, ph, dh, kh, sh), _out_extent(W, pw, dw, kw, sw)
    out = torch.empty(B, Co, Do, Ho, Wo, dtype=torch.float32, device=input.device)
    ws = Workspace.get(input.device, lib.dlka_deform_conv_pack3d_workspace_bytes(
        B, C, D, H, W, Co, kd, kh, kw, sd, sh, sw, pd, ph, pw, dd, dh, dw, groups, deformable_groups))
    with torch.cuda.device(input.device):
        st = lib.dlka_deform_conv_pack3d_forward(
            dptr(input, "input"), dptr(offset_weight.contiguous(), "conv_offset.weight"),
            dptr(offset_bias.contiguous(), "conv_offset.bias"), dptr(weight.contiguous(), "weight"),
            dptr(bias.contiguous(), "bias"), dptr(out), B, C, D, H, W, Co, kd, kh, kw, sd, sh, sw, pd, ph, pw,
            dd, dh, dw, groups, deformable_groups, im2col_step, _math(math), ws.data_ptr(), ws.numel(),
            stream_ptr(input.device))
    check(st, "dlka_deform_conv_pack3d_forward")
    return out


def deform_conv_pack2d(input, offset_weight, offset_bias, weight, bias, stride, padding, dilation, math=None) -> torch.Tensor:
    """DeformConv.forward (2D/deformable_LKA/deformable_LKA.py:27-30): offset_net + deformable conv in one call."""
    if not input.is_cuda:
        raise RuntimeError("Not implemented on the CPU (deformablelka_b200 is CUDA-only)")
    if input.dim() != 4:
        raise RuntimeError("expected a 4-D NCHW tensor")
    input = input.contiguous()
    sh, sw = _pair(stride); ph, pw = _pair(padding); dh, dw = _pair(dilation)
    B, C, H, W = input.shape
    Co, Cg, kh, kw = weight.shape
    groups = C // Cg
    Ho, Wo = _out_extent(H, ph, dh, kh, sh), _out_extent(W, pw, dw, kw, sw)
    out = torch.empty(B, Co, Ho, Wo, dtype=torch.float32, device=input.device)
    ws = Workspace.get(input.device, lib.dlka_deform_conv_pack2d_workspace_bytes(B, C, H, W, Co, kh, kw, sh, sw, ph, pw, dh, dw, groups))
    with torch.cuda.device(input.device):
        st = lib.dlka_deform_conv_pack2d_forward(
            dptr(input, "input"), dptr(offset_weight.contiguous(), "offset_net.weight"),
            dptr(offset_bias.contiguous(), "offset_net.bias"), dptr(weight.contiguous(), "weight"),
            dptr(None if bias is None else bias.contiguous(), "bias"), dptr(out),
            B, C, H, W, Co, kh, kw, sh, sw, ph, pw, dh, dw, groups, _math(math), ws.data_ptr(), ws.numel(),
            stream_ptr(input.device))
    check(st, "dlka_deform_conv_pack2d_forward")
    return out


# ----------------------------------------------------------------------------------------------
# blocks
# ----------------------------------------------------------------------------------------------
def _params_struct(cls, tensors: dict):
    keep = []
    s = cls()
    for name, _ in cls._fields_:
        t = tensors.get(name)
        if name == "dw_geom":   # ((k0), (dil0), (k1), (dil1)) or None = the synapse network's shapes
            if t is not None:
                g = DwGeom3d()
                for dst, src in zip((g.conv0_k, g.conv0_dil, g.conv_spatial_k, g.conv_spatial_dil), t):
                    dst[0], dst[1], dst[2] = (int(v) for v in src)
                keep.append(g)
                s.dw_geom = ctypes.pointer(g)
            continue
        if t is None:
            setattr(s, name, None)
            continue
        t = t.detach()
        if not t.is_contiguous():
            t = t.contiguous()
        keep.append(t)
        setattr(s, name, dptr(t, name))
    return s, keep


def lka3d_deform_forward(params: dict, x: torch.Tensor, math=None) -> torch.Tensor:
    """LKA3d_deform.forward on NCDHW input (transformerblock.py:644-652)."""
    if x.dim() != 5:
        raise RuntimeError("expected a 5-D NCDHW tensor")
    x = x.contiguous()
    B, C, D1, D2, D3 = x.shape
    y = torch.empty_like(x)
    s, keep = _params_struct(Block3dParams, params)
    ws = Workspace.get(x.device, lib.dlka_lka3d_deform_workspace_bytes(B, C, D1, D2, D3))
    with torch.cuda.device(x.device):
        st = lib.dlka_lka3d_deform_forward(ctypes.byref(s), dptr(x, "x"), dptr(y), B, C, D1, D2, D3, _math(math),
                                           ws.data_ptr(), ws.numel(), stream_ptr(x.device))
    check(st, "dlka_lka3d_deform_forward")
    return y


def _pack_signature(params: dict):
    """Identity of the parameter VALUES as far as the host can see it: storage pointer + in-place version counter of every tensor."""
    return tuple((k, t.data_ptr(), t._version) for k, t in sorted(params.items()) if torch.is_tensor(t))


def _packed_slot(cache, key, params, nbytes, device):
    """(buffer, valid) of the prepacked weights for this (module, shape): valid when the buffer was filled from the same
    parameter values (dlka_*_forward_packed).  `cache` is a plain dict owned by the module."""
    sig = _pack_signature(params)
    ent = cache.get(key)
    if ent is None or ent[0].numel() < nbytes or ent[0].device != device:
        ent = [torch.empty(nbytes, dtype=torch.uint8, device=device), None]
        cache[key] = ent
    return ent, sig


def lka_attention3d_deform_forward(params: dict, x: torch.Tensor, B: int, C: int, H: int, W: int, D: int,
                                   math=None, cache: Optional[dict] = None) -> torch.Tensor:
    """LKA_Attention3d_deform.forward on tokens [B, N, C], N = H*W*D (transformerblock.py:664-673).
    cache: a dict owned by the calling module -> the packed weights are kept there and re-packed only when a parameter changed."""
    if x.dim() != 3 or x.shape[0] != B or x.shape[1] != H * W * D or x.shape[2] != C:
        raise RuntimeError(f"expected tokens of shape {(B, H * W * D, C)}, got {tuple(x.shape)}")
    x = x.contiguous()
    y = torch.empty_like(x)
    s, keep = _params_struct(Block3dParams, params)
    ws = Workspace.get(x.device, lib.dlka_lka_attention3d_deform_workspace_bytes(B, C, H, W, D))
    if cache is not None and all(t.is_contiguous() for t in params.values() if torch.is_tensor(t)):
        mm = _math(math)
        ent, sig = _packed_slot(cache, ("attn3d", B, C, H, W, D, mm), params, lib.dlka_lka_attention3d_deform_packed_bytes(C), x.device)
        with torch.cuda.device(x.device):
            st = lib.dlka_lka_attention3d_deform_forward_packed(
                ctypes.byref(s), dptr(x, "x"), dptr(y), B, C, H, W, D, mm, ent[0].data_ptr(), ent[0].numel(),
                1 if ent[1] == sig else 0, ws.data_ptr(), ws.numel(), stream_ptr(x.device))
        check(st, "dlka_lka_attention3d_deform_forward_packed")
        ent[1] = sig
        return y
    with torch.cuda.device(x.device):
        st = lib.dlka_lka_attention3d_deform_forward(ctypes.byref(s), dptr(x, "x"), dptr(y), B, C, H, W, D, _math(math),
                                                     ws.data_ptr(), ws.numel(), stream_ptr(x.device))
    check(st, "dlka_lka_attention3d_deform_forward")
    return y


def lka_attention3d_deform_forward_host(params: dict, x_host: torch.Tensor, y_host: torch.Tensor, B, C, H, W, D,
                                        device, math=None) -> torch.Tensor:
    """Host-buffer variant: x_host / y_host are CPU (ideally pinned) fp32 tensors; parameters live on `device`."""
    assert x_host.device.type == "cpu" and y_host.device.type == "cpu" and x_host.is_contiguous() and y_host.is_contiguous()
    device = torch.device(device)
    n = B * H * W * D * C
    s, keep = _params_struct(Block3dParams, params)
    dev_scratch = _HostScratch.get(device, 2 * n * 4)
    ws = Workspace.get(device, lib.dlka_lka_attention3d_deform_workspace_bytes(B, C, H, W, D))
    with torch.cuda.device(device):
        st = lib.dlka_lka_attention3d_deform_forward_host(
            ctypes.byref(s), x_host.data_ptr(), y_host.data_ptr(), B, C, H, W, D, _math(math),
            dev_scratch.data_ptr(), dev_scratch.numel(), ws.data_ptr(), ws.numel(), stream_ptr(device))
    check(st, "dlka_lka_attention3d_deform_forward_host")
    return y_host


def deformable_lka_block2d_forward(attn_params: dict, block: dict, x: torch.Tensor, H: int, W: int, hidden: int,
                                   eps1: float, eps2: float, math=None) -> torch.Tensor:
    """deformableLKABlock.forward on tokens [B, N, C], N = H*W (MaxViT_deform_LKA.py:165-189)."""
    if x.dim() != 3 or x.shape[1] != H * W:
        raise RuntimeError(f"expected tokens of shape [B, {H * W}, C], got {tuple(x.shape)}")
    x = x.contiguous()
    B, N, C = x.shape
    y = torch.empty_like(x)
    keep = []
    s = _lib.LkaBlock2dParams()
    a, k = _params_struct(Block2dParams, attn_params)
    keep.append(k)
    s.attn = a
    for name in ("norm1_weight", "norm1_bias", "layer_scale_1", "norm2_weight", "norm2_bias", "fc1_weight", "fc1_bias",
                 "dw_weight", "dw_bias", "fc2_weight", "fc2_bias", "layer_scale_2"):
        t = block[name].detach().contiguous()
        keep.append(t)
        setattr(s, name, dptr(t, name))
    s.eps1, s.eps2, s.hidden = float(eps1), float(eps2), int(hidden)
    ws = Workspace.get(x.device, lib.dlka_deformable_lka_block2d_workspace_bytes(B, C, H, W, hidden))
    with torch.cuda.device(x.device):
        st = lib.dlka_deformable_lka_block2d_forward(ctypes.byref(s), dptr(x, "x"), dptr(y), B, C, H, W, _math(math),
                                                     ws.data_ptr(), ws.numel(), stream_ptr(x.device))
    check(st, "dlka_deformable_lka_block2d_forward")
    return y


def lka_transformer3d_prenorm_forward(attn_params: dict, norm_weight, norm_bias, eps, gamma, pos_embed, x, B, C, H, W, D,
                                      math=None) -> torch.Tensor:
    """x' = x + pos_embed; y = x' + gamma * LKA_Attention3d_deform(LayerNorm(x'))  (transformerblock.py:620-624)."""
    if x.dim() != 3 or tuple(x.shape) != (B, H * W * D, C):
        raise RuntimeError(f"expected tokens of shape {(B, H * W * D, C)}, got {tuple(x.shape)}")
    x = x.contiguous()
    y = torch.empty_like(x)
    s, keep = _params_struct(Block3dParams, attn_params)
    nw, nb, gm = norm_weight.detach().contiguous(), norm_bias.detach().contiguous(), gamma.detach().contiguous()
    pe = None if pos_embed is None else pos_embed.detach().reshape(-1, C).contiguous()
    if pe is not None and pe.shape[0] != H * W * D:
        raise RuntimeError(f"pos_embed must have {H * W * D} rows, got {pe.shape[0]}")
    ws = Workspace.get(x.device, lib.dlka_lka_transformer3d_prenorm_workspace_bytes(B, C, H, W, D))
    with torch.cuda.device(x.device):
        st = lib.dlka_lka_transformer3d_prenorm_forward(ctypes.byref(s), dptr(nw, "norm.weight"), dptr(nb, "norm.bias"),
                                                        ctypes.c_float(float(eps)), dptr(gm, "gamma"), dptr(pe, "pos_embed"),
                                                        dptr(x, "x"), dptr(y), B, C, H, W, D, _math(math), ws.data_ptr(),
                                                        ws.numel(), stream_ptr(x.device))
    check(st, "dlka_lka_transformer3d_prenorm_forward")
    return y


def lka_transformer3d_block_forward(attn_params: dict, tail: dict, eps: float, slope: float, x: torch.Tensor, B, C, H, W, D,
                                    math=None) -> torch.Tensor:
    """Whole TransformerBlock_3D_single_deform_LKA on tokens (transformerblock.py:617-630, inference mode).
    `tail` holds norm_weight, norm_bias, gamma, pos_embed (or None), conv1_weight, bn1_scale, bn1_shift, conv2_weight,
    bn2_scale, bn2_shift, conv8_weight, conv8_bias."""
    if x.dim() != 3 or tuple(x.shape) != (B, H * W * D, C):
        raise RuntimeError(f"expected tokens of shape {(B, H * W * D, C)}, got {tuple(x.shape)}")
    x = x.contiguous()
    y = torch.empty_like(x)
    keep = []
    s = _lib.Transformer3dParams()
    a, k = _params_struct(Block3dParams, attn_params)
    keep.append(k)
    s.attn = a
    for name, _ in _lib.Transformer3dParams._fields_:
        if name in ("attn", "eps", "lrelu_slope"):
            continue
        t = tail.get(name)
        if t is None:
            setattr(s, name, None)
            continue
        t = t.detach().reshape(-1, C).contiguous() if name == "pos_embed" else t.detach().contiguous()
        keep.append(t)
        setattr(s, name, dptr(t, name))
    s.eps, s.lrelu_slope = float(eps), float(slope)
    ws = Workspace.get(x.device, lib.dlka_lka_transformer3d_block_workspace_bytes(B, C, H, W, D))
    with torch.cuda.device(x.device):
        st = lib.dlka_lka_transformer3d_block_forward(ctypes.byref(s), dptr(x, "x"), dptr(y), B, C, H, W, D, _math(math),
                                                      ws.data_ptr(), ws.numel(), stream_ptr(x.device))
    check(st, "dlka_lka_transformer3d_block_forward")
    return y


_NUMA_POLICY = {"off": 0, "local": 1, "interleave": 2}


def host_numa_policy(policy=None) -> int:
    """Placement of pinned host buffers: "local" (the device's NUMA node, default), "interleave" (all nodes) or "off";
    the DLKA_HOST_NUMA environment variable overrides the default."""
    import os
    name = policy if policy is not None else os.environ.get("DLKA_HOST_NUMA", "local")
    try:
        return _NUMA_POLICY[str(name).lower()]
    except KeyError:
        raise ValueError(f"unknown host NUMA policy {name!r}; expected one of {sorted(_NUMA_POLICY)}")


def bind_host_thread(device) -> int:
    """Pin the calling host thread to the cores of `device`'s NUMA node (and prefer that node for its allocations).
    Returns the node, or -1 when the topology is unknown (nothing changed)."""
    device = torch.device(device)
    return int(lib.dlka_host_bind_thread(device.index if device.index is not None else torch.cuda.current_device()))


def pinned_empty(shape, device, policy=None, write_combined: bool = False) -> torch.Tensor:
    """A page-locked fp32 host tensor placed for `device` (dlka_host_alloc: mbind before first touch + cudaHostRegister).
    write_combined=True: an INPUT buffer the host only writes (cudaHostAllocWriteCombined; CPU reads from it are very slow).
    The memory is released when the tensor (and every view of it) is garbage-collected."""
    import weakref
    device = torch.device(device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    n = 1
    for d in shape:
        n *= int(d)
    ptr = ctypes.c_void_p()
    with torch.cuda.device(idx):
        check(lib.dlka_host_alloc(ctypes.byref(ptr), max(n, 1) * 4, idx, host_numa_policy(policy) | (16 if write_combined else 0)),
              "dlka_host_alloc")
    buf = (ctypes.c_float * max(n, 1)).from_address(ptr.value)
    t = torch.frombuffer(buf, dtype=torch.float32, count=n).view(*shape)
    weakref.finalize(buf, lib.dlka_host_free, ptr)   # `buf` is kept alive by the tensor's storage
    return t


class HostPipe:
    """Streaming host-buffer pipeline (dlka_host_pipe_*): keeps `depth` steps in flight so that the H2D copy of the next
    step and the D2H copy of the previous one overlap the compute of the current step.

    Lifetime rules enforced here (without ever blocking the host in submit): every step's host tensors and contiguous
    parameter copies are held until the library reports that step's last D2H copy as finished (dlka_host_pipe_completed,
    polled non-blocking) or until wait(); the device staging buffer is never regrown while steps are in flight."""

    def __init__(self, device, depth: int = 2):
        self.device = torch.device(device)
        self.depth = depth
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(lib.dlka_host_pipe_create(ctypes.byref(self._h), depth), "dlka_host_pipe_create")
        self._scratch = None
        self._inflight = []     # [(step index, references)]
        self._step = 0

    def _release_finished(self) -> None:
        if not self._inflight:
            return
        done = int(lib.dlka_host_pipe_completed(self._h))
        if done < 0:
            check(done, "dlka_host_pipe_completed")
        self._inflight = [(i, r) for i, r in self._inflight if i >= done]

    def submit(self, params: dict, x_host: torch.Tensor, y_host: torch.Tensor, B, C, H, W, D, math=None) -> None:
        assert x_host.device.type == "cpu" and y_host.device.type == "cpu" and x_host.is_contiguous() and y_host.is_contiguous()
        n = B * H * W * D * C
        need = self.depth * 2 * n * 4
        if self._scratch is None or self._scratch.numel() < need:
            self.wait()                               # copies on the library's private streams may still use the old buffer
            self._scratch = None
            self._scratch = torch.empty(need, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            self._release_finished()
        s, keep = _params_struct(Block3dParams, params)
        ws = Workspace.get(self.device, lib.dlka_lka_attention3d_deform_workspace_bytes(1, C, H, W, D))
        with torch.cuda.device(self.device):
            st = lib.dlka_lka_attention3d_deform_forward_host_async(
                self._h, ctypes.byref(s), x_host.data_ptr(), y_host.data_ptr(), B, C, H, W, D, _math(math),
                self._scratch.data_ptr(), self._scratch.numel(), ws.data_ptr(), ws.numel(), stream_ptr(self.device))
        self._inflight.append((self._step, [keep, x_host, y_host, ws]))
        check(st, "dlka_lka_attention3d_deform_forward_host_async")
        self._step += 1

    def join(self) -> None:
        """Order the current CUDA stream after every result copy enqueued so far (no host sync)."""
        with torch.cuda.device(self.device):
            check(lib.dlka_host_pipe_join(self._h, stream_ptr(self.device)), "dlka_host_pipe_join")

    def wait(self) -> None:
        with torch.cuda.device(self.device):
            torch.cuda.current_stream(self.device).synchronize()
            check(lib.dlka_host_pipe_wait(self._h), "dlka_host_pipe_wait")
        self._inflight = []

    def __del__(self):
        try:
            if self._h:
                lib.dlka_host_pipe_destroy(self._h)   # synchronises the private streams first
                self._h = ctypes.c_void_p()
        except Exception:
            pass


class _HostScratch:
    _bufs = {}

    @classmethod
    def get(cls, device, nbytes):
        key = device.index if device.index is not None else torch.cuda.current_device()
        b = cls._bufs.get(key)
        if b is None or b.numel() < nbytes:
            cls._bufs[key] = None
            b = torch.empty(nbytes, dtype=torch.uint8, device=device)
            cls._bufs[key] = b
        return b


def deformable_lka2d_forward(params: dict, x: torch.Tensor, math=None) -> torch.Tensor:
    """deformable_LKA.forward on NCHW input (deformable_LKA.py:98-104)."""
    if x.dim() != 4:
        raise RuntimeError("expected a 4-D NCHW tensor")
    x = x.contiguous()
    B, C, H, W = x.shape
    y = torch.empty_like(x)
    s, keep = _params_struct(Block2dParams, params)
    ws = Workspace.get(x.device, lib.dlka_deformable_lka2d_workspace_bytes(B, C, H, W))
    with torch.cuda.device(x.device):
        st = lib.dlka_deformable_lka2d_forward(ctypes.byref(s), dptr(x, "x"), dptr(y), B, C, H, W, _math(math),
                                               ws.data_ptr(), ws.numel(), stream_ptr(x.device))
    check(st, "dlka_deformable_lka2d_forward")
    return y


def deformable_lka_attention2d_forward(params: dict, x: torch.Tensor, math=None, cache: Optional[dict] = None) -> torch.Tensor:
    """deformable_LKA_Attention.forward on NCHW input (deformable_LKA.py:133-140); `cache` as in lka_attention3d_deform_forward."""
    if x.dim() != 4:
        raise RuntimeError("expected a 4-D NCHW tensor")
    x = x.contiguous()
    B, C, H, W = x.shape
    y = torch.empty_like(x)
    s, keep = _params_struct(Block2dParams, params)
    ws = Workspace.get(x.device, lib.dlka_deformable_lka_attention2d_workspace_bytes(B, C, H, W))
    if cache is not None and all(t.is_contiguous() for t in params.values() if torch.is_tensor(t)):
        mm = _math(math)
        ent, sig = _packed_slot(cache, ("attn2d", B, C, H, W, mm), params, lib.dlka_deformable_lka_attention2d_packed_bytes(C), x.device)
        with torch.cuda.device(x.device):
            st = lib.dlka_deformable_lka_attention2d_forward_packed(
                ctypes.byref(s), dptr(x, "x"), dptr(y), B, C, H, W, mm, ent[0].data_ptr(), ent[0].numel(),
                1 if ent[1] == sig else 0, ws.data_ptr(), ws.numel(), stream_ptr(x.device))
        check(st, "dlka_deformable_lka_attention2d_forward_packed")
        ent[1] = sig
        return y
    with torch.cuda.device(x.device):
        st = lib.dlka_deformable_lka_attention2d_forward(ctypes.byref(s), dptr(x, "x"), dptr(y), B, C, H, W, _math(math),
                                                         ws.data_ptr(), ws.numel(), stream_ptr(x.device))
    check(st, "dlka_deformable_lka_attention2d_forward")
    return y


def deform_conv3d_backward(input: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, offset: torch.Tensor,
                           grad_output: torch.Tensor, kernel_size, stride, padding, dilation, group: int,
                           deformable_group: int, im2col_step: int = 64, math=None):
    """Same argument meaning and error behaviour as ``D3D.deform_conv_backward`` (deform_conv_cuda.cu:128-285); returns
    (grad_input, grad_offset, grad_weight, grad_bias)."""
    if not input.is_cuda:
        raise RuntimeError("Not implemented on the CPU")  # deform_conv.h:84
    if not input.is_contiguous():
        raise RuntimeError("input tensor has to be contiguous")  # deform_conv_cuda.cu:150
    if not weight.is_contiguous():
        raise RuntimeError("weight tensor has to be contiguous")  # :151
    kd, kh, kw = _triple(kernel_size)
    sd, sh, sw = _triple(stride)
    pd, ph, pw = _triple(padding)
    dd, dh, dw = _triple(dilation)
    B, C, D, H, W = input.shape
    Co, Cg, kd_, kh_, kw_ = weight.shape
    if (kd_, kh_, kw_) != (kd, kh, kw):
        raise RuntimeError(f"Input shape and kernel shape wont match: ({kd} x {kh} x {kw} vs {kd_} x {kh_} x {kw_}).")  # :186
    if C != Cg * group:
        raise RuntimeError(f"Input shape and kernel channels wont match: ({C} vs {Cg * group}).")  # :189
    step = min(B, im2col_step)
    if B % step:
        raise RuntimeError(f"batch({B}) must divide im2col_step({step})")  # :178
    Do, Ho, Wo = _out_extent(D, pd, dd, kd, sd), _out_extent(H, ph, dh, kh, sh), _out_extent(W, pw, dw, kw, sw)
    if grad_output.shape[0] != B:
        raise RuntimeError(f"Input shape and grad_out batch wont match: ({B} vs {grad_output.shape[0]}).")  # :196
    if grad_output.shape[1] != Co:
        raise RuntimeError(f"Input shape and grad_out channels_out wont match: ({Co} vs {grad_output.shape[1]}).")  # :199
    if tuple(grad_output.shape[2:]) != (Do, Ho, Wo):
        raise RuntimeError(f"Input shape and grad_out shape wont match: ({Do} x {Ho} x {Wo} vs "
                           f"{grad_output.shape[2]} x {grad_output.shape[3]} x {grad_output.shape[4]}).")  # :202
    offset = offset.contiguous()
    grad_output = grad_output.contiguous()
    gi, go = torch.empty_like(input), torch.empty_like(offset)
    gw, gb = torch.empty_like(weight), torch.empty(Co, dtype=torch.float32, device=input.device)
    ws = Workspace.get(input.device, lib.dlka_deform_conv3d_backward_workspace_bytes(
        B, C, D, H, W, Co, kd, kh, kw, sd, sh, sw, pd, ph, pw, dd, dh, dw, group, deformable_group))
    with torch.cuda.device(input.device):
        st = lib.dlka_deform_conv3d_backward(
            dptr(input, "input"), dptr(weight, "weight"), dptr(offset, "offset"), dptr(grad_output, "grad_output"),
            dptr(gi), dptr(go), dptr(gw), dptr(gb), B, C, D, H, W, Co, kd, kh, kw, sd, sh, sw, pd, ph, pw, dd, dh, dw,
            group, deformable_group, im2col_step, _math(math), ws.data_ptr(), ws.numel(), stream_ptr(input.device))
    check(st, "dlka_deform_conv3d_backward")
    return gi, go, gw, gb


def linear_tokens_forward(x: torch.Tensor, weight: torch.Tensor, bias=None, add=None, math=None) -> torch.Tensor:
    """nn.Linear on tokens [..., K] -> [..., N], optionally + add (MyDecoderLayer.x1_linear and its skip add,
    2D/networks/MaxViT_deform_LKA.py:604-607)."""
    if not x.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    if x.shape[-1] != weight.shape[1]:
        raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied ({tuple(x.shape)} and {tuple(weight.t().shape)})")
    x = x.contiguous()
    K, N = weight.shape[1], weight.shape[0]
    M = x.numel() // K
    y = torch.empty(*x.shape[:-1], N, device=x.device, dtype=torch.float32)
    w = weight.detach().contiguous()
    b = bias.detach().contiguous() if bias is not None else None
    if add is not None:
        add = add.contiguous()
        if add.numel() != M * N:
            raise RuntimeError(f"skip tensor of {add.numel()} elements does not match the [{M}, {N}] output")
    ws = Workspace.get(x.device, lib.dlka_linear_tokens_workspace_bytes(K, N))
    with torch.cuda.device(x.device):
        st = lib.dlka_linear_tokens_forward(dptr(x, "x"), dptr(w, "weight"), dptr(b) if b is not None else None,
                                            dptr(add) if add is not None else None, dptr(y), M, K, N, _math(math),
                                            ws.data_ptr(), ws.numel(), stream_ptr(x.device))
    check(st, "dlka_linear_tokens_forward")
    return y


def patch_expand2d_forward(x: torch.Tensor, expand_weight, norm_weight, norm_bias, eps: float, H: int, W: int, scale: int,
                           math=None) -> torch.Tensor:
    """PatchExpand.forward (scale 2) / FinalPatchExpand_X4.forward (scale 4) on tokens [B, H*W, dim]
    (2D/networks/MaxViT_deform_LKA.py:488-545)."""
    if not x.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    B, L, dim = x.shape
    assert L == H * W, "input feature has wrong size"   # MaxViT_deform_LKA.py:506,536
    x = x.contiguous()
    cg = dim // 2 if scale == 2 else dim
    y = torch.empty(B, scale * scale * H * W, cg, device=x.device, dtype=torch.float32)
    w, g, b = (t.detach().contiguous() for t in (expand_weight, norm_weight, norm_bias))
    ws = Workspace.get(x.device, lib.dlka_patch_expand2d_workspace_bytes(B, H, W, dim, scale))
    with torch.cuda.device(x.device):
        st = lib.dlka_patch_expand2d_forward(dptr(x, "x"), dptr(w, "expand.weight"), dptr(g), dptr(b), float(eps), dptr(y),
                                             B, H, W, dim, scale, _math(math), ws.data_ptr(), ws.numel(), stream_ptr(x.device))
    check(st, "dlka_patch_expand2d_forward")
    return y

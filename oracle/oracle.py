"""CPU oracle for the D-LKA hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(``deformablelka_b200``) never does.

The oracle restates, on CPU, what the reference computes on the hot path:

* stock layers (``nn.Conv2d/Conv3d``, ``nn.GELU``) are the very PyTorch CPU ops the
  reference calls (2D/deformable_LKA/deformable_LKA.py:10-16,95,128-131;
  3D/d_lka_former/network_architecture/synapse/transformerblock.py:637-641,659-662);
* the 3D deformable convolution (CUDA-only in the reference: 3D/dcn/src/deform_conv.h:46)
  is restated in C (``dlka_oracle.c``: im2col per deform_im2col_cuda.cuh:192-265,
  sampler per :26-72) followed by ``torch.addmm`` exactly as deform_conv_cuda.cu:113-123;
* the 2D deformable convolution is ``torchvision.ops.deform_conv2d`` on CPU (the
  reference's own third-party dependency, 2D/deformable_LKA/deformable_LKA.py:18-25)
  with a C restatement beside it that the tests pin against torchvision.

Module classes below keep the reference's attribute names so state_dicts are
interchangeable with the reference modules and with ``deformablelka_b200``.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from typing import Optional, Sequence

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libdlka_oracle.so")
_SRC = os.path.join(_HERE, "dlka_oracle.c")


def build(force: bool = False) -> str:
    """Compile dlka_oracle.c with gcc (OpenMP) into oracle/libdlka_oracle.so."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        cmd = ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-shared", "-fPIC", "-o", _SO, _SRC, "-lm"]
        subprocess.check_call(cmd)
    return _SO


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _fp(t: Optional[torch.Tensor]):
    if t is None:
        return ctypes.c_void_p(0)
    assert t.dtype == torch.float32 and t.is_contiguous() and t.device.type == "cpu"
    return ctypes.c_void_p(t.data_ptr())


def _ip(t: torch.Tensor):
    assert t.dtype == torch.int32 and t.is_contiguous()
    return ctypes.c_void_p(t.data_ptr())


def _triple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v, v)


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def out_extent(n, pad, dil, k, stride):
    return (n + 2 * pad - (dil * (k - 1) + 1)) // stride + 1


# --------------------------------------------------------------------------------------
# 3D deformable convolution (D3D.deform_conv_forward restated)
# --------------------------------------------------------------------------------------
def deform_conv3d(input: torch.Tensor, offset: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor,
                  stride=1, padding=0, dilation=1, groups: int = 1, deformable_groups: int = 1,
                  im2col_step: int = 64, chunk: int = 32768) -> torch.Tensor:
    """Follows deform_conv_cuda_forward (3D/dcn/src/cuda/deform_conv_cuda.cu:18-126).

    ``im2col_step`` only changes how the reference batches its im2col buffer; the result does
    not depend on it, but the reference's divisibility assert (cu:61-63) is kept.
    """
    input = input.contiguous().float()
    offset = offset.contiguous().float()
    weight = weight.contiguous().float()
    B, C, D, H, W = input.shape
    Co, Cg, kd, kh, kw = weight.shape
    sd, sh, sw = _triple(stride)
    pd, ph, pw = _triple(padding)
    dd, dh, dw = _triple(dilation)
    if C != Cg * groups:
        raise RuntimeError(f"Input shape and kernel channels wont match: ({C} vs {Cg * groups}).")
    if C % groups or Co % groups:
        raise RuntimeError("channels and channels_out must divide group")
    step = min(B, im2col_step)
    if B % step:
        raise RuntimeError(f"batch({B}) must divide im2col_step({step})")
    K = kd * kh * kw
    Do, Ho, Wo = out_extent(D, pd, dd, kd, sd), out_extent(H, ph, dh, kh, sh), out_extent(W, pw, dw, kw, sw)
    if tuple(offset.shape) != (B, deformable_groups * 3 * K, Do, Ho, Wo):
        raise RuntimeError(f"offset shape {tuple(offset.shape)} != {(B, deformable_groups * 3 * K, Do, Ho, Wo)}")
    Vo = Do * Ho * Wo
    out = torch.empty(B, Co, Vo, dtype=torch.float32)
    w_g = weight.view(groups, Co // groups, Cg * K)
    b_g = bias.float().view(groups, Co // groups)
    L = lib()
    for b in range(B):
        for v0 in range(0, Vo, chunk):
            v1 = min(Vo, v0 + chunk)
            cols = torch.empty(C * K, v1 - v0, dtype=torch.float32)
            L.oracle_deform_im2col3d(_fp(input), _fp(offset), _fp(cols), B, C, D, H, W, kd, kh, kw,
                                     sd, sh, sw, pd, ph, pw, dd, dh, dw, deformable_groups, b,
                                     ctypes.c_int64(v0), ctypes.c_int64(v1))
            cols_g = cols.view(groups, Cg * K, v1 - v0)
            for g in range(groups):
                # at::addmm(bias_g, columns_g^T, weight_g^T)   (cu:113-119)
                o = torch.addmm(b_g[g], cols_g[g].t(), w_g[g].t())  # [n, Co/g]
                out[b, g * (Co // groups):(g + 1) * (Co // groups), v0:v1] = o.t()
    return out.view(B, Co, Do, Ho, Wo)


def deform_conv3d_c(input, offset, weight, bias, stride=1, padding=0, dilation=1, groups=1, deformable_groups=1):
    """All-C variant (naive GEMM) -- small shapes only; cross-checks the addmm variant."""
    input = input.contiguous().float(); offset = offset.contiguous().float(); weight = weight.contiguous().float()
    B, C, D, H, W = input.shape
    Co, Cg, kd, kh, kw = weight.shape
    sd, sh, sw = _triple(stride); pd, ph, pw = _triple(padding); dd, dh, dw = _triple(dilation)
    Do, Ho, Wo = out_extent(D, pd, dd, kd, sd), out_extent(H, ph, dh, kh, sh), out_extent(W, pw, dw, kw, sw)
    out = torch.empty(B, Co, Do, Ho, Wo, dtype=torch.float32)
    lib().oracle_deform_conv3d_forward(_fp(input), _fp(offset), _fp(weight), _fp(bias.contiguous().float()), _fp(out),
                                       B, C, D, H, W, Co, kd, kh, kw, sd, sh, sw, pd, ph, pw, dd, dh, dw,
                                       groups, deformable_groups)
    return out


def sample_indices3d(offset, in_size, kernel_size, stride=1, padding=0, dilation=1, deformable_groups=1):
    """Integer planes of the sampler (floor per axis + validity / corner mask), for bit-exact parity."""
    offset = offset.contiguous().float()
    D, H, W = in_size
    kd, kh, kw = _triple(kernel_size)
    sd, sh, sw = _triple(stride); pd, ph, pw = _triple(padding); dd, dh, dw = _triple(dilation)
    B = offset.shape[0]
    K = kd * kh * kw
    Vo = offset.shape[2] * offset.shape[3] * offset.shape[4]
    low = torch.empty(B * deformable_groups, Vo, K, 3, dtype=torch.int32)
    mask = torch.empty(B * deformable_groups, Vo, K, dtype=torch.int32)
    lib().oracle_sample_indices3d(_fp(offset), _ip(low), _ip(mask), B, D, H, W, kd, kh, kw, sd, sh, sw,
                                  pd, ph, pw, dd, dh, dw, deformable_groups)
    return low, mask


# --------------------------------------------------------------------------------------
# 2D deformable convolution
# --------------------------------------------------------------------------------------
def deform_conv2d_c(input, offset, weight, bias=None, stride=1, padding=0, dilation=1, mask=None):
    """C restatement of torchvision.ops.deform_conv2d (same argument meaning)."""
    input = input.contiguous().float(); offset = offset.contiguous().float(); weight = weight.contiguous().float()
    B, C, H, W = input.shape
    Co, Cg, kh, kw = weight.shape
    sh, sw = _pair(stride); ph, pw = _pair(padding); dh, dw = _pair(dilation)
    n_off = offset.shape[1] // (2 * kh * kw)
    n_wg = C // Cg
    Ho, Wo = out_extent(H, ph, dh, kh, sh), out_extent(W, pw, dw, kw, sw)
    out = torch.empty(B, Co, Ho, Wo, dtype=torch.float32)
    lib().oracle_deform_conv2d_forward(_fp(input), _fp(offset), _fp(None if mask is None else mask.contiguous().float()),
                                       _fp(weight), _fp(None if bias is None else bias.contiguous().float()), _fp(out),
                                       B, C, H, W, Co, kh, kw, sh, sw, ph, pw, dh, dw, n_wg, n_off)
    return out


def deform_conv2d(input, offset, weight, bias=None, stride=1, padding=0, dilation=1, mask=None):
    """The reference's own dependency on CPU: torchvision.ops.deform_conv2d."""
    import torchvision
    return torchvision.ops.deform_conv2d(input, offset, weight, bias, _pair(stride), _pair(padding), _pair(dilation), mask)


def sample_indices2d(offset, in_size, kernel_size, stride=1, padding=0, dilation=1, n_offset_grps=1):
    offset = offset.contiguous().float()
    H, W = in_size
    kh, kw = _pair(kernel_size)
    sh, sw = _pair(stride); ph, pw = _pair(padding); dh, dw = _pair(dilation)
    B = offset.shape[0]
    K = kh * kw
    P = offset.shape[2] * offset.shape[3]
    low = torch.empty(B * n_offset_grps, P, K, 2, dtype=torch.int32)
    mask = torch.empty(B * n_offset_grps, P, K, dtype=torch.int32)
    lib().oracle_sample_indices2d(_fp(offset), _ip(low), _ip(mask), B, H, W, kh, kw, sh, sw, ph, pw, dh, dw, n_offset_grps)
    return low, mask


# --------------------------------------------------------------------------------------
# Module restatements (same attribute names / state_dict keys as the reference)
# --------------------------------------------------------------------------------------
class _DeformConv2dParams(nn.Module):
    """Holds ``weight`` like torchvision.ops.DeformConv2d(bias=False) (deformable_LKA.py:18-25)."""

    def __init__(self, in_channels, out_channels, kernel_size, padding, groups, stride, dilation):
        super().__init__()
        self.kernel_size = _pair(kernel_size); self.padding = _pair(padding)
        self.stride = _pair(stride); self.dilation = _pair(dilation); self.groups = groups
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels // groups, *self.kernel_size))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        self.register_parameter("bias", None)


class DeformConv2D(nn.Module):
    """Restates ``DeformConv`` (2D/deformable_LKA/deformable_LKA.py:5-30)."""

    def __init__(self, in_channels, groups, kernel_size=(3, 3), padding=1, stride=1, dilation=1, bias=True, impl="torchvision"):
        super().__init__()
        self.offset_net = nn.Conv2d(in_channels, 2 * kernel_size[0] * kernel_size[1], kernel_size=kernel_size,
                                    padding=padding, stride=stride, dilation=dilation, bias=True)
        self.deform_conv = _DeformConv2dParams(in_channels, in_channels, kernel_size, padding, groups, stride, dilation)
        self.impl = impl

    def forward(self, x):
        offsets = self.offset_net(x)
        dc = self.deform_conv
        fn = deform_conv2d if self.impl == "torchvision" else deform_conv2d_c
        return fn(x, offsets, dc.weight, None, dc.stride, dc.padding, dc.dilation, None)


class deformable_LKA(nn.Module):
    """Restates deformable_LKA (2D/deformable_LKA/deformable_LKA.py:90-104)."""

    def __init__(self, dim, impl="torchvision"):
        super().__init__()
        self.conv0 = DeformConv2D(dim, kernel_size=(5, 5), padding=2, groups=dim, impl=impl)
        self.conv_spatial = DeformConv2D(dim, kernel_size=(7, 7), stride=1, padding=9, groups=dim, dilation=3, impl=impl)
        self.conv1 = nn.Conv2d(dim, dim, 1)

    def forward(self, x):
        u = x.clone()
        attn = self.conv0(x)
        attn = self.conv_spatial(attn)
        attn = self.conv1(attn)
        return u * attn


class deformable_LKA_Attention(nn.Module):
    """Restates deformable_LKA_Attention (2D/deformable_LKA/deformable_LKA.py:124-140)."""

    def __init__(self, d_model, impl="torchvision"):
        super().__init__()
        self.proj_1 = nn.Conv2d(d_model, d_model, 1)
        self.activation = nn.GELU()
        self.spatial_gating_unit = deformable_LKA(d_model, impl=impl)
        self.proj_2 = nn.Conv2d(d_model, d_model, 1)

    def forward(self, x):
        shortcut = x.clone()
        x = self.proj_1(x)
        x = self.activation(x)
        x = self.spatial_gating_unit(x)
        x = self.proj_2(x)
        return x + shortcut


class DeformConv3D(nn.Module):
    """Restates DeformConv (3D/dcn/modules/deform_conv.py:15-63)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation=1, groups=1,
                 deformable_groups=1, im2col_step=64, bias=True):
        super().__init__()
        if in_channels % groups != 0:
            raise ValueError('in_channels {} must be divisible by groups {}'.format(in_channels, groups))
        if out_channels % groups != 0:
            raise ValueError('out_channels {} must be divisible by groups {}'.format(out_channels, groups))
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride = _triple(kernel_size), _triple(stride)
        self.padding, self.dilation = _triple(padding), _triple(dilation)
        self.groups, self.deformable_groups, self.im2col_step = groups, deformable_groups, im2col_step
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels // groups, *self.kernel_size))
        self.bias = nn.Parameter(torch.empty(out_channels))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        fan_in, _ = nn.init._calculate_fan_in_and_fan_out(self.weight)
        bound = 1 / math.sqrt(fan_in)
        nn.init.uniform_(self.bias, -bound, bound)
        if not bias:
            self.bias.requires_grad = False

    def forward(self, input, offset):
        K = self.kernel_size[0] * self.kernel_size[1] * self.kernel_size[2]
        assert 3 * self.deformable_groups * K == offset.shape[1]
        return deform_conv3d(input, offset, self.weight, self.bias, self.stride, self.padding, self.dilation,
                             self.groups, self.deformable_groups, self.im2col_step)


class DeformConvPack3D(DeformConv3D):
    """Restates DeformConvPack (synapse/deform_conv.py:67-105): conv_offset ignores dilation, zero-init."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation=1, groups=1,
                 deformable_groups=1, im2col_step=64, bias=True, lr_mult=0.1):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups,
                         deformable_groups, im2col_step, bias)
        oc = self.deformable_groups * 3 * self.kernel_size[0] * self.kernel_size[1] * self.kernel_size[2]
        self.conv_offset = nn.Conv3d(self.in_channels, oc, kernel_size=self.kernel_size, stride=self.stride,
                                     padding=self.padding, bias=True)
        self.conv_offset.weight.data.zero_()
        self.conv_offset.bias.data.zero_()

    def forward(self, input):
        offset = self.conv_offset(input)
        return deform_conv3d(input, offset, self.weight, self.bias, self.stride, self.padding, self.dilation,
                             self.groups, self.deformable_groups, self.im2col_step)


class LKA3d_deform(nn.Module):
    """Restates LKA3d_deform (synapse/transformerblock.py:634-652)."""

    def __init__(self, dim):
        super().__init__()
        self.conv0 = nn.Conv3d(dim, dim, 5, padding=2, groups=dim)
        self.conv_spatial = nn.Conv3d(dim, dim, 7, stride=1, padding=9, groups=dim, dilation=3)
        self.deform_conv = DeformConvPack3D(in_channels=dim, out_channels=dim, kernel_size=(3, 3, 3), stride=1, padding=1)
        self.conv1 = nn.Conv3d(dim, dim, 1)

    def forward(self, x):
        u = x.clone()
        attn = self.conv0(x)
        attn = self.conv_spatial(attn)
        attn = attn.contiguous()
        attn = self.deform_conv(attn)
        attn = self.conv1(attn)
        return u * attn


class LKA_Attention3d_deform(nn.Module):
    """Restates LKA_Attention3d_deform (synapse/transformerblock.py:655-673)."""

    def __init__(self, d_model):
        super().__init__()
        self.proj_1 = nn.Conv3d(d_model, d_model, 1)
        self.activation = nn.GELU()
        self.spatial_gating_unit = LKA3d_deform(d_model)
        self.proj_2 = nn.Conv3d(d_model, d_model, 1)

    def forward(self, x, B, C, H, W, D):
        x = x.permute(0, 2, 1).reshape(B, C, H, W, D)
        shortcut = x.clone()
        x = self.proj_1(x)
        x = self.activation(x)
        x = self.spatial_gating_unit(x)
        x = self.proj_2(x)
        x = x + shortcut
        return x.reshape(B, C, H * W * D).permute(0, 2, 1)


class LKA3d_deform_ACDC(LKA3d_deform):
    """Restates the ACDC variant (acdc/transformerblock.py:210-253): same forward, depthwise stencil shapes per dim."""

    def __init__(self, dim):
        nn.Module.__init__(self)
        if dim == 32 or dim == 64:
            kernel_dwd, dilation_dwd, padding_dwd, kernel_dw, padding_dw = (5, 7, 7), (3, 3, 3), (6, 9, 9), 5, 2
        elif dim == 128:
            kernel_dwd, dilation_dwd, padding_dwd, kernel_dw, padding_dw = (3, 5, 5), (1, 3, 3), (1, 6, 6), 5, 2
        elif dim == 256:
            kernel_dwd, dilation_dwd, padding_dwd, kernel_dw, padding_dw = 3, 1, 1, 3, 1
        else:
            raise ValueError("Unknown dim: {}".format(dim))
        self.conv0 = nn.Conv3d(dim, dim, kernel_size=kernel_dw, padding=padding_dw, groups=dim)
        self.conv_spatial = nn.Conv3d(dim, dim, kernel_size=kernel_dwd, stride=1, padding=padding_dwd, groups=dim,
                                      dilation=dilation_dwd)
        self.conv1 = nn.Conv3d(dim, dim, 1)
        self.deform_conv = DeformConvPack3D(in_channels=dim, out_channels=dim, kernel_size=(3, 3, 3), stride=1, padding=1)


class LKA_Attention3d_deform_ACDC(LKA_Attention3d_deform):
    """Restates the ACDC attention wrapper (acdc/transformerblock.py:255-275)."""

    def __init__(self, d_model):
        nn.Module.__init__(self)
        self.proj_1 = nn.Conv3d(d_model, d_model, 1)
        self.activation = nn.GELU()
        self.spatial_gating_unit = LKA3d_deform_ACDC(d_model)
        self.proj_2 = nn.Conv3d(d_model, d_model, 1)


def deform_conv3d_autograd(x, offset, weight, bias, stride=(1, 1, 1), padding=(1, 1, 1), dilation=(1, 1, 1), group=1,
                           deformable_group=1):
    """Differentiable pure-torch restatement of the 3D deformable conv forward (weight groups: deform_conv_cuda.cu:84-110,
    deformable groups: the offset channel ((dg * K + tap) * 3 + axis) of deform_im2col_cuda.cuh:222-235), used as the
    backward oracle: autograd through it yields exactly the gradients dmcn_get_gradient_weight / dmcn_get_coordinate_weight
    define (deform_im2col_cuda.cuh:74-190) -- the corner weights are differentiated, floor() has zero derivative, and the
    sample / corner validity rules are those of dmcn_im2col_bilinear (cuh:30-65, 248).  Small shapes only."""
    B, C, D, H, W = x.shape
    Co, _, kd, kh, kw = weight.shape
    sd, sh, sw = stride
    pd, ph, pw = padding
    dd, dh, dw = dilation
    Do = (D + 2 * pd - (dd * (kd - 1) + 1)) // sd + 1
    Ho = (H + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1
    Wo = (W + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1
    K = kd * kh * kw
    od = (torch.arange(Do, dtype=x.dtype) * sd - pd).view(1, Do, 1, 1)
    oh = (torch.arange(Ho, dtype=x.dtype) * sh - ph).view(1, 1, Ho, 1)
    ow = (torch.arange(Wo, dtype=x.dtype) * sw - pw).view(1, 1, 1, Wo)
    cpd = C // deformable_group
    gcols = []
    for dgi in range(deformable_group):
      xf = x[:, dgi * cpd:(dgi + 1) * cpd].reshape(B, cpd, D * H * W)
      cols = []
      for tap in range(K):
        i, j, k = tap // (kh * kw), (tap // kw) % kh, tap % kw
        oc = 3 * (dgi * K + tap)
        p_d = od + i * dd + offset[:, oc]               # [B, Do, Ho, Wo]
        p_h = oh + j * dh + offset[:, oc + 1]
        p_w = ow + k * dw + offset[:, oc + 2]
        valid = (p_d > -1) & (p_h > -1) & (p_w > -1) & (p_d < D) & (p_h < H) & (p_w < W)
        d0, h0, w0 = torch.floor(p_d).detach(), torch.floor(p_h).detach(), torch.floor(p_w).detach()
        ld, lh, lw = p_d - d0, p_h - h0, p_w - w0
        val = 0
        for cd in (0, 1):
            for ch in (0, 1):
                for cw in (0, 1):
                    di, hi, wi = d0 + cd, h0 + ch, w0 + cw
                    ok = valid & (di >= 0) & (di <= D - 1) & (hi >= 0) & (hi <= H - 1) & (wi >= 0) & (wi <= W - 1)
                    wgt = (ld if cd else 1 - ld) * (lh if ch else 1 - lh) * (lw if cw else 1 - lw)
                    idx = (di.clamp(0, D - 1) * H + hi.clamp(0, H - 1)) * W + wi.clamp(0, W - 1)
                    g = torch.gather(xf, 2, idx.long().view(B, 1, -1).expand(B, cpd, -1)).view(B, cpd, Do, Ho, Wo)
                    val = val + g * (wgt * ok.to(x.dtype)).unsqueeze(1)
        cols.append(val)
      gcols.append(torch.stack(cols, dim=2))             # [B, cpd, K, Do, Ho, Wo]
    col = torch.cat(gcols, dim=1)                        # [B, C, K, Do, Ho, Wo]
    cpg, copg = C // group, Co // group
    outs = [torch.einsum("bckdhw,ock->bodhw", col[:, gi * cpg:(gi + 1) * cpg],
                         weight[gi * copg:(gi + 1) * copg].reshape(copg, cpg, K)) for gi in range(group)]
    return torch.cat(outs, dim=1) + bias.view(1, Co, 1, 1, 1)


class PatchExpand(nn.Module):
    """Restates PatchExpand (2D/networks/MaxViT_deform_LKA.py:488-513); the einops rearrange is spelled out with view/permute."""

    def __init__(self, input_resolution, dim, dim_scale=2, norm_layer=nn.LayerNorm):
        super().__init__()
        self.input_resolution, self.dim = input_resolution, dim
        self.expand = nn.Linear(dim, 2 * dim, bias=False) if dim_scale == 2 else nn.Identity()
        self.norm = norm_layer(dim // dim_scale)

    @staticmethod
    def _shuffle(x, B, H, W, p, c):   # "b h w (p1 p2 c) -> b (h p1) (w p2) c"
        return x.view(B, H, W, p, p, c).permute(0, 1, 3, 2, 4, 5).reshape(B, H * p, W * p, c)

    def forward(self, x):
        H, W = self.input_resolution
        x = self.expand(x)
        B, L, C = x.shape
        assert L == H * W, "input feature has wrong size"
        x = self._shuffle(x, B, H, W, 2, C // 4).reshape(B, -1, C // 4)
        return self.norm(x.clone())


class FinalPatchExpand_X4(nn.Module):
    """Restates FinalPatchExpand_X4 (:516-545)."""

    def __init__(self, input_resolution, dim, dim_scale=4, norm_layer=nn.LayerNorm):
        super().__init__()
        self.input_resolution, self.dim, self.dim_scale = input_resolution, dim, dim_scale
        self.expand = nn.Linear(dim, 16 * dim, bias=False)
        self.output_dim = dim
        self.norm = norm_layer(self.output_dim)

    def forward(self, x):
        H, W = self.input_resolution
        x = self.expand(x)
        B, L, C = x.shape
        assert L == H * W, "input feature has wrong size"
        x = PatchExpand._shuffle(x, B, H, W, self.dim_scale, C // (self.dim_scale ** 2)).reshape(B, -1, self.output_dim)
        return self.norm(x.clone())


class MyDecoderLayer(nn.Module):
    """Restates MyDecoderLayer (:548-620) from the oracle's own blocks."""

    def __init__(self, input_size, in_out_chan, head_count, token_mlp_mode, n_class=9, norm_layer=nn.LayerNorm, is_last=False):
        super().__init__()
        out_dim, x1_dim = in_out_chan[1], in_out_chan[4]
        self.x1_linear = nn.Linear(x1_dim, out_dim)
        if not is_last:
            self.layer_up = PatchExpand(input_resolution=input_size, dim=out_dim, dim_scale=2, norm_layer=norm_layer)
            self.last_layer = None
        else:
            self.layer_up = FinalPatchExpand_X4(input_resolution=input_size, dim=out_dim, dim_scale=4, norm_layer=norm_layer)
            self.last_layer = nn.Conv2d(out_dim, n_class, 1)
        self.layer_lka_1 = deformableLKABlock(dim=out_dim)
        self.layer_lka_2 = deformableLKABlock(dim=out_dim)

    def forward(self, x1, x2=None):
        if x2 is None:
            return self.layer_up(x1)
        b, h, w, c = x2.shape
        x2 = x2.view(b, -1, c)
        cat_linear_x = self.x1_linear(x1) + x2
        t = self.layer_lka_2(self.layer_lka_1(cat_linear_x, h, w), h, w)
        if self.last_layer:
            return self.last_layer(self.layer_up(t).view(b, 4 * h, 4 * w, -1).permute(0, 3, 1, 2))
        return self.layer_up(t)


def sliding_window_predict_oracle(network, x, patch_size, num_classes, step_size=0.5, do_mirroring=True, mirror_axes=(0, 1, 2),
                                  use_gaussian=True):
    """numpy restatement of SegmentationNetwork._internal_predict_3D_3Dconv_tiled in its default (host aggregation, fp32)
    mode (3D/d_lka_former/network_architecture/neural_network.py:292-428), with _compute_steps_for_sliding_window (:267-290),
    _get_gaussian (:251-264) and _internal_maybe_mirror_and_pred_3D (:502-556) spelled out the way the reference writes them
    (explicit loops, the eight mirror cases).  pad_nd_image is batchgenerators' (not in the reference tree): symmetric
    constant padding, remainder above.  Returns (segmentation, class_probabilities)."""
    import numpy as np
    from scipy.ndimage import gaussian_filter
    x = np.asarray(x, dtype=np.float32)
    assert len(x.shape) == 4
    new_shape = [max(a, b) for a, b in zip(x.shape[1:], patch_size)]
    diff = [n - o for n, o in zip(new_shape, x.shape[1:])]
    below = [d // 2 for d in diff]
    above = [d // 2 + d % 2 for d in diff]
    data = np.pad(x, [(0, 0)] + [(b, a) for b, a in zip(below, above)], mode="constant")
    slicer = [slice(b, b + o) for b, o in zip(below, x.shape[1:])]
    image_size = data.shape[1:]
    target = [i * step_size for i in patch_size]
    num_steps = [int(np.ceil((i - k) / j)) + 1 for i, j, k in zip(image_size, target, patch_size)]
    steps = []
    for dim in range(3):
        max_step_value = image_size[dim] - patch_size[dim]
        actual = max_step_value / (num_steps[dim] - 1) if num_steps[dim] > 1 else 99999999999
        steps.append([int(np.round(actual * i)) for i in range(num_steps[dim])])
    num_tiles = len(steps[0]) * len(steps[1]) * len(steps[2])
    gaussian = None
    if use_gaussian and num_tiles > 1:
        tmp = np.zeros(patch_size)
        tmp[tuple(i // 2 for i in patch_size)] = 1
        gaussian = gaussian_filter(tmp, [i * (1. / 8) for i in patch_size], 0, mode="constant", cval=0)
        gaussian = (gaussian / np.max(gaussian) * 1).astype(np.float32)
        gaussian[gaussian == 0] = np.min(gaussian[gaussian != 0])
    add = gaussian if gaussian is not None else np.ones(patch_size, dtype=np.float32)
    agg = np.zeros([num_classes] + list(data.shape[1:]), dtype=np.float32)
    nb = np.zeros([num_classes] + list(data.shape[1:]), dtype=np.float32)

    def pred_one(patch):
        xt = torch.from_numpy(np.ascontiguousarray(patch))
        res = torch.zeros([1, num_classes] + list(xt.shape[2:]))
        n = 2 ** len(mirror_axes) if do_mirroring else 1
        cases = [((), True), ((4,), 2 in mirror_axes), ((3,), 1 in mirror_axes), ((4, 3), 2 in mirror_axes and 1 in mirror_axes),
                 ((2,), 0 in mirror_axes), ((4, 2), 0 in mirror_axes and 2 in mirror_axes),
                 ((3, 2), 0 in mirror_axes and 1 in mirror_axes),
                 ((4, 3, 2), 0 in mirror_axes and 1 in mirror_axes and 2 in mirror_axes)]
        for m, (dims, on) in enumerate(cases[:8 if do_mirroring else 1]):
            if not on:
                continue
            with torch.no_grad():
                p = torch.softmax(network(torch.flip(xt, dims) if dims else xt), 1)
            res += 1 / n * (torch.flip(p, dims) if dims else p)
        if gaussian is not None:
            res[:, :] *= torch.from_numpy(gaussian)
        return res[0].numpy()

    for lx in steps[0]:
        for ly in steps[1]:
            for lz in steps[2]:
                ux, uy, uz = lx + patch_size[0], ly + patch_size[1], lz + patch_size[2]
                agg[:, lx:ux, ly:uy, lz:uz] += pred_one(data[None, :, lx:ux, ly:uy, lz:uz])
                nb[:, lx:ux, ly:uy, lz:uz] += add
    sl = tuple([slice(None)] + slicer)
    probs = agg[sl] / nb[sl]
    return probs.argmax(0), probs


def randomize_offsets_(module: nn.Module, std: float = 0.05, bias_range: float = 1.0, seed: int = 0) -> None:
    """re-initialise the zero-initialised 3D ``conv_offset`` so offsets are non-trivial."""
    g = torch.Generator().manual_seed(seed)
    for m in module.modules():
        if hasattr(m, "conv_offset"):
            with torch.no_grad():
                m.conv_offset.weight.copy_(torch.randn(m.conv_offset.weight.shape, generator=g) * std)
                m.conv_offset.bias.copy_((torch.rand(m.conv_offset.bias.shape, generator=g) * 2 - 1) * bias_range)


# --------------------------------------------------------------------------------------
# Row N1: the enclosing transformer blocks, restated from stock PyTorch layers
# --------------------------------------------------------------------------------------
class DWConvLKA(nn.Module):
    """Restates DWConvLKA (2D/networks/MaxViT_deform_LKA.py:18-27)."""

    def __init__(self, dim=768):
        super().__init__()
        self.dwconv = nn.Conv2d(dim, dim, 3, 1, 1, bias=True, groups=dim)

    def forward(self, x):
        return self.dwconv(x)


class Mlp(nn.Module):
    """Restates Mlp (MaxViT_deform_LKA.py:29-52), linear=False, drop=0."""

    def __init__(self, in_features, hidden_features=None, out_features=None):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Conv2d(in_features, hidden_features, 1)
        self.dwconv = DWConvLKA(hidden_features)
        self.act = nn.GELU()
        self.fc2 = nn.Conv2d(hidden_features, out_features, 1)

    def forward(self, x):
        return self.fc2(self.act(self.dwconv(self.fc1(x))))


class deformableLKABlock(nn.Module):
    """Restates deformableLKABlock (MaxViT_deform_LKA.py:142-189), eval mode (drop_path = Identity)."""

    def __init__(self, dim, mlp_ratio=4., impl="torchvision"):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn = deformable_LKA_Attention(dim, impl=impl)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio))
        self.layer_scale_1 = nn.Parameter(1e-2 * torch.ones((dim)), requires_grad=True)
        self.layer_scale_2 = nn.Parameter(1e-2 * torch.ones((dim)), requires_grad=True)

    def forward(self, x, H, W):
        B, N, C = x.shape
        x = x.permute(0, 2, 1).view(B, C, H, W)
        y = self.norm1(x.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)
        y = self.attn(y)
        x = x + self.layer_scale_1.unsqueeze(-1).unsqueeze(-1) * y
        y = self.norm2(x.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)
        y = self.mlp(y)
        x = x + self.layer_scale_2.unsqueeze(-1).unsqueeze(-1) * y
        return x.view(B, C, N).permute(0, 2, 1)


def transformer3d_attention_half(norm: nn.LayerNorm, gamma: torch.Tensor, epa_block: LKA_Attention3d_deform,
                                 pos_embed: Optional[torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """Restates TransformerBlock_3D_single_deform_LKA.forward up to the gamma residual (transformerblock.py:617-624):
    x [B,C,H,W,D] -> tokens; (+pos_embed); attn = x + gamma * epa_block(norm(x), B, C, H, W, D).  Returns tokens."""
    B, C, H, W, D = x.shape
    t = x.reshape(B, C, H * W * D).permute(0, 2, 1)
    if pos_embed is not None:
        t = t + pos_embed
    return t + gamma * epa_block(norm(t), B, C, H, W, D)


class UnetResBlock3D(nn.Module):
    """Restates monai/dynunet UnetResBlock(3, C, C, kernel_size=3, stride=1, norm_name="batch")
    (3D/d_lka_former/network_architecture/dynunet_block.py:12-80) from stock layers; conv layers keep monai's
    ``.conv`` nesting so state_dict keys match (conv1.conv.weight, norm1.*, ...)."""

    class _Conv(nn.Module):
        def __init__(self, c):
            super().__init__()
            self.conv = nn.Conv3d(c, c, 3, stride=1, padding=1, bias=False)

        def forward(self, x):
            return self.conv(x)

    def __init__(self, c):
        super().__init__()
        self.conv1 = UnetResBlock3D._Conv(c)
        self.conv2 = UnetResBlock3D._Conv(c)
        self.lrelu = nn.LeakyReLU(negative_slope=0.01, inplace=False)
        self.norm1 = nn.BatchNorm3d(c)
        self.norm2 = nn.BatchNorm3d(c)

    def forward(self, inp):
        out = self.lrelu(self.norm1(self.conv1(inp)))
        out = self.norm2(self.conv2(out))
        return self.lrelu(out + inp)


def transformer3d_block(norm, gamma, epa_block, pos_embed, conv51: UnetResBlock3D, conv8_conv: nn.Conv3d, x: torch.Tensor):
    """Restates TransformerBlock_3D_single_deform_LKA.forward (transformerblock.py:617-630), eval mode
    (Dropout3d(0.1) in conv8 is the identity).  x [B,C,H,W,D] -> [B,C,H,W,D]."""
    B, C, H, W, D = x.shape
    attn = transformer3d_attention_half(norm, gamma, epa_block, pos_embed, x)
    attn_skip = attn.reshape(B, H, W, D, C).permute(0, 4, 1, 2, 3)
    return attn_skip + conv8_conv(conv51(attn_skip))

"""2D drop-in modules: same class names, constructor arguments, attribute names and state_dict keys as
the reference's 2D/deformable_LKA/deformable_LKA.py, with forward passes executed by libdlka_b200
(hand-written sm_100a CUDA behind the C ABI of include/dlka.h).

    reference                                   here
    DeformConv                (:5-30)           DeformConv
    torchvision.ops.DeformConv2d (:18-25)       DeformConv2d   (parameter holder + operator call)
    deformable_LKA            (:90-104)         deformable_LKA
    deformable_LKA_Attention  (:124-140)        deformable_LKA_Attention
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
from torch.nn.modules.utils import _pair

from . import ops
from .lka3d import needs_autograd


class DeformConv2dFunction(torch.autograd.Function):
    """Autograd bridge of the 2D operator: library forward (dlka_deform_conv2d_forward) and library backward
    (dlka_deform_conv2d_backward) -- the gradients torch.ops.torchvision.deform_conv2d's autograd returns."""

    @staticmethod
    def forward(ctx, input, offset, weight, bias, mask, stride, padding, dilation):
        ctx.cfg = (stride, padding, dilation)
        ctx.save_for_backward(input, offset, weight, bias, mask)
        return ops.deform_conv2d(input, offset, weight, bias, stride, padding, dilation, mask)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_output):
        input, offset, weight, bias, mask = ctx.saved_tensors
        stride, padding, dilation = ctx.cfg
        gi, go, gw, gm, gb = ops.deform_conv2d_backward(input, offset, weight, mask, grad_output, stride, padding, dilation,
                                                        need_bias_grad=bias is not None)
        return gi, go, gw, gb, gm, None, None, None


def deform_conv2d_autograd(input, offset, weight, bias=None, stride=(1, 1), padding=(0, 0), dilation=(1, 1), mask=None):
    """``ops.deform_conv2d`` that autograd can differentiate (used by the modules below when a gradient can be asked for)."""
    return DeformConv2dFunction.apply(input, offset, weight, bias, mask, _pair(stride), _pair(padding), _pair(dilation))


class DeformConv2d(nn.Module):
    """Mirror of ``torchvision.ops.DeformConv2d`` (same ctor / parameters / forward signature)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True):
        super().__init__()
        if in_channels % groups != 0:
            raise ValueError("in_channels must be divisible by groups")
        if out_channels % groups != 0:
            raise ValueError("out_channels must be divisible by groups")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride = _pair(kernel_size), _pair(stride)
        self.padding, self.dilation, self.groups = _pair(padding), _pair(dilation), groups
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels // groups, *self.kernel_size))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self) -> None:
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = nn.init._calculate_fan_in_and_fan_out(self.weight)
            bound = 1 / math.sqrt(fan_in)
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, input, offset, mask=None):
        if needs_autograd(self, input, offset, mask):
            return deform_conv2d_autograd(input, offset, self.weight, self.bias, self.stride, self.padding, self.dilation, mask)
        return ops.deform_conv2d(input, offset, self.weight, self.bias, self.stride, self.padding, self.dilation, mask)


class DeformConv(nn.Module):
    """offset_net (dense conv, same k/pad/stride/dil) + depthwise deformable conv (deformable_LKA.py:5-30)."""

    def __init__(self, in_channels, groups, kernel_size=(3, 3), padding=1, stride=1, dilation=1, bias=True):
        super().__init__()
        self.offset_net = nn.Conv2d(in_channels=in_channels, out_channels=2 * kernel_size[0] * kernel_size[1],
                                    kernel_size=kernel_size, padding=padding, stride=stride, dilation=dilation, bias=True)
        self.deform_conv = DeformConv2d(in_channels=in_channels, out_channels=in_channels, kernel_size=kernel_size,
                                        padding=padding, groups=groups, stride=stride, dilation=dilation, bias=False)

    def forward(self, x):
        if needs_autograd(self, x):
            # training: the reference's own two steps (deformable_LKA.py:27-30) -- stock offset_net, differentiable operator
            with torch.backends.cudnn.flags(allow_tf32=False):   # offsets are sampling positions: keep them fp32
                offsets = self.offset_net(x)
            return self.deform_conv(x, offsets)
        return ops.deform_conv_pack2d(x, self.offset_net.weight, self.offset_net.bias, self.deform_conv.weight,
                                      self.deform_conv.bias, self.deform_conv.stride, self.deform_conv.padding,
                                      self.deform_conv.dilation)


def _block2d_params(lka: "deformable_LKA", attn=None) -> dict:
    p = {
        "conv0_offset_weight": lka.conv0.offset_net.weight, "conv0_offset_bias": lka.conv0.offset_net.bias,
        "conv0_deform_weight": lka.conv0.deform_conv.weight,
        "conv_spatial_offset_weight": lka.conv_spatial.offset_net.weight,
        "conv_spatial_offset_bias": lka.conv_spatial.offset_net.bias,
        "conv_spatial_deform_weight": lka.conv_spatial.deform_conv.weight,
        "conv1_weight": lka.conv1.weight, "conv1_bias": lka.conv1.bias,
    }
    if attn is not None:
        p.update({"proj_1_weight": attn.proj_1.weight, "proj_1_bias": attn.proj_1.bias,
                  "proj_2_weight": attn.proj_2.weight, "proj_2_bias": attn.proj_2.bias})
    return p


class deformable_LKA(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.conv0 = DeformConv(dim, kernel_size=(5, 5), padding=2, groups=dim)
        self.conv_spatial = DeformConv(dim, kernel_size=(7, 7), stride=1, padding=9, groups=dim, dilation=3)
        self.conv1 = nn.Conv2d(dim, dim, 1)

    def forward(self, x):
        if needs_autograd(self, x):   # differentiable composition (deformable_LKA.py:98-104)
            return x * self.conv1(self.conv_spatial(self.conv0(x)))
        # u * conv1(conv_spatial(conv0(x)))  in one library call
        return ops.deformable_lka2d_forward(_block2d_params(self), x)


class deformable_LKA_Attention(nn.Module):
    def __init__(self, d_model):
        super().__init__()
        self.proj_1 = nn.Conv2d(d_model, d_model, 1)
        self.activation = nn.GELU()
        self.spatial_gating_unit = deformable_LKA(d_model)
        self.proj_2 = nn.Conv2d(d_model, d_model, 1)

    def forward(self, x):
        if needs_autograd(self, x):   # differentiable composition (deformable_LKA.py:133-140)
            return self.proj_2(self.spatial_gating_unit(self.activation(self.proj_1(x)))) + x
        # proj_1 -> GELU -> gating unit -> proj_2 -> + shortcut in one library call
        return ops.deformable_lka_attention2d_forward(_block2d_params(self.spatial_gating_unit, self), x,
                                                      cache=self.__dict__.setdefault("_dlka_pack_cache", {}))

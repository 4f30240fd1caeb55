"""3D operator boundary: mirrors 3D/dcn/functions/deform_conv_func.py and 3D/dcn/modules/deform_conv.py
(identical copies: 3D/d_lka_former/network_architecture/synapse/{deform_conv_func,deform_conv}.py).

``D3D.deform_conv_forward`` / ``D3D.deform_conv_backward`` are replaced by ``ops.deform_conv3d_forward`` /
``ops.deform_conv3d_backward`` (C ABI ``dlka_deform_conv3d_forward`` / ``dlka_deform_conv3d_backward``; the backward is
SURVEY.md 8f row N2, groups = deformable groups = 1).  Inference (no grad) takes the fused one-call path.
"""
from __future__ import annotations

import math

import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn import init
from torch.nn.modules.utils import _triple

from . import ops


class DeformConvFunction(Function):
    @staticmethod
    def forward(ctx, input, offset, weight, bias, stride, padding, dilation, group, deformable_groups, im2col_step):
        ctx.stride = _triple(stride)
        ctx.padding = _triple(padding)
        ctx.dilation = _triple(dilation)
        ctx.kernel_size = _triple(weight.shape[2:5])
        ctx.group = group
        ctx.deformable_groups = deformable_groups
        ctx.im2col_step = im2col_step
        output = ops.deform_conv3d_forward(input, weight, bias, offset, ctx.kernel_size, ctx.stride, ctx.padding,
                                           ctx.dilation, ctx.group, ctx.deformable_groups, ctx.im2col_step)
        ctx.save_for_backward(input, offset, weight, bias)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        input, offset, weight, bias = ctx.saved_tensors
        grad_input, grad_offset, grad_weight, grad_bias = ops.deform_conv3d_backward(
            input, weight, bias, offset, grad_output, ctx.kernel_size, ctx.stride, ctx.padding, ctx.dilation, ctx.group,
            ctx.deformable_groups, ctx.im2col_step)
        return grad_input, grad_offset, grad_weight, grad_bias, None, None, None, None, None, None


class DeformConv(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation=1, groups=1,
                 deformable_groups=1, im2col_step=64, bias=True):
        super().__init__()
        if in_channels % groups != 0:
            raise ValueError('in_channels {} must be divisible by groups {}'.format(in_channels, groups))
        if out_channels % groups != 0:
            raise ValueError('out_channels {} must be divisible by groups {}'.format(out_channels, groups))
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _triple(kernel_size)
        self.stride = _triple(stride)
        self.padding = _triple(padding)
        self.dilation = _triple(dilation)
        self.groups = groups
        self.deformable_groups = deformable_groups
        self.im2col_step = im2col_step
        self.use_bias = bias
        self.weight = nn.Parameter(torch.Tensor(out_channels, in_channels // groups, *self.kernel_size))
        self.bias = nn.Parameter(torch.Tensor(out_channels))
        self.reset_parameters()
        if not self.use_bias:
            self.bias.requires_grad = False

    def reset_parameters(self):
        init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = init._calculate_fan_in_and_fan_out(self.weight)
            bound = 1 / math.sqrt(fan_in)
            init.uniform_(self.bias, -bound, bound)

    def forward(self, input, offset):
        assert 3 * self.deformable_groups * self.kernel_size[0] * self.kernel_size[1] * self.kernel_size[2] == \
            offset.shape[1]
        return DeformConvFunction.apply(input, offset, self.weight, self.bias, self.stride, self.padding,
                                        self.dilation, self.groups, self.deformable_groups, self.im2col_step)


_DeformConv = DeformConvFunction.apply


class DeformConvPack(DeformConv):
    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation=1, groups=1,
                 deformable_groups=1, im2col_step=64, bias=True, lr_mult=0.1):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups,
                         deformable_groups, im2col_step, bias)
        out_channels = self.deformable_groups * 3 * self.kernel_size[0] * self.kernel_size[1] * self.kernel_size[2]
        # the reference passes no dilation to conv_offset (synapse/deform_conv.py:80-85)
        self.conv_offset = nn.Conv3d(self.in_channels, out_channels, kernel_size=self.kernel_size,
                                     stride=self.stride, padding=self.padding, bias=True)
        self.conv_offset.lr_mult = lr_mult
        self.init_offset()

    def init_offset(self):
        self.conv_offset.weight.data.zero_()
        self.conv_offset.bias.data.zero_()

    def forward(self, input):
        if torch.is_grad_enabled() and (input.requires_grad or any(p.requires_grad for p in self.parameters())):
            # training: the reference's own two steps (synapse/deform_conv.py:93-105) so autograd sees conv_offset (stock
            # nn.Conv3d) and DeformConvFunction (library forward + dlka_deform_conv3d_backward)
            with torch.backends.cudnn.flags(allow_tf32=False):   # offsets are sampling positions: keep them fp32
                offset = self.conv_offset(input)
            return DeformConvFunction.apply(input.contiguous(), offset, self.weight, self.bias, self.stride, self.padding,
                                            self.dilation, self.groups, self.deformable_groups, self.im2col_step)
        return ops.deform_conv_pack3d(input, self.conv_offset.weight, self.conv_offset.bias, self.weight, self.bias,
                                      self.stride, self.padding, self.dilation, self.groups, self.deformable_groups,
                                      self.im2col_step)

"""ACDC variant of the 3D D-LKA block: same modules, depthwise stencil shapes chosen per channel count
(3D/d_lka_former/network_architecture/acdc/transformerblock.py:210-275).  Class names, ctor arguments and state_dict
keys are the reference's; each forward is ONE call into libdlka_b200 (the stencil shapes travel in dlkaDwGeom3d)."""
from __future__ import annotations

import torch.nn as nn

from . import lka3d
from .deform_conv3d import DeformConvPack


def _triple(v):
    return (v, v, v) if isinstance(v, int) else tuple(v)


def stencil_shapes(dim: int):
    """(kernel_dw, padding_dw, kernel_dwd, dilation_dwd, padding_dwd) exactly as acdc/transformerblock.py:214-236."""
    if dim in (32, 64):
        return 5, 2, (5, 7, 7), (3, 3, 3), (6, 9, 9)
    if dim == 128:
        return 5, 2, (3, 5, 5), (1, 3, 3), (1, 6, 6)
    if dim == 256:
        return 3, 1, 3, 1, 1
    raise ValueError("Unknown dim: {}".format(dim))


class LKA3d_deform(lka3d.LKA3d_deform):
    def __init__(self, dim):
        nn.Module.__init__(self)
        kernel_dw, padding_dw, kernel_dwd, dilation_dwd, padding_dwd = stencil_shapes(dim)
        self.conv0 = nn.Conv3d(dim, dim, kernel_size=kernel_dw, padding=padding_dw, groups=dim)
        self.conv_spatial = nn.Conv3d(dim, dim, kernel_size=kernel_dwd, stride=1, padding=padding_dwd, groups=dim,
                                      dilation=dilation_dwd)
        self.conv1 = nn.Conv3d(dim, dim, 1)
        self.deform_conv = DeformConvPack(in_channels=dim, out_channels=dim, kernel_size=(3, 3, 3), stride=1, padding=1)
        self.dw_geom = (_triple(kernel_dw), (1, 1, 1), _triple(kernel_dwd), _triple(dilation_dwd))


class LKA_Attention3d_deform(lka3d.LKA_Attention3d_deform):
    def __init__(self, d_model):
        nn.Module.__init__(self)
        self.proj_1 = nn.Conv3d(d_model, d_model, 1)
        self.activation = nn.GELU()
        self.spatial_gating_unit = LKA3d_deform(d_model)
        self.proj_2 = nn.Conv3d(d_model, d_model, 1)


def _transformer_block():
    from .blocks import TransformerBlock_3D_single_deform_LKA as _Base

    class TransformerBlock_3D_single_deform_LKA(_Base):
        """acdc/transformerblock.py:146-207 -- the same block around the ACDC attention (same names, ctor, keys; the whole
        forward is still one library call, the stencil shapes travel in dlkaDwGeom3d)."""

        @staticmethod
        def _attention_class():
            return LKA_Attention3d_deform

    return TransformerBlock_3D_single_deform_LKA


def __getattr__(name):   # lazy: blocks.py imports lka3d, which must not import this module back at import time
    if name == "TransformerBlock_3D_single_deform_LKA":
        cls = _transformer_block()
        globals()[name] = cls
        return cls
    raise AttributeError(name)

"""3D drop-in modules: ``LKA3d_deform`` and ``LKA_Attention3d_deform`` with the reference's names, ctor
arguments and state_dict keys (3D/d_lka_former/network_architecture/synapse/transformerblock.py:634-673).
Each forward is ONE call into libdlka_b200."""
from __future__ import annotations

import torch.nn as nn

from . import ops
from .deform_conv3d import DeformConvPack


def _block3d_params(lka: "LKA3d_deform", attn=None) -> dict:
    dc = lka.deform_conv
    p = {
        "conv0_weight": lka.conv0.weight, "conv0_bias": lka.conv0.bias,
        "conv_spatial_weight": lka.conv_spatial.weight, "conv_spatial_bias": lka.conv_spatial.bias,
        "conv_offset_weight": dc.conv_offset.weight, "conv_offset_bias": dc.conv_offset.bias,
        "deform_weight": dc.weight, "deform_bias": dc.bias,
        "conv1_weight": lka.conv1.weight, "conv1_bias": lka.conv1.bias,
    }
    if attn is not None:
        p.update({"proj_1_weight": attn.proj_1.weight, "proj_1_bias": attn.proj_1.bias,
                  "proj_2_weight": attn.proj_2.weight, "proj_2_bias": attn.proj_2.bias})
    geom = getattr(lka, "dw_geom", None)
    if geom is not None:
        p["dw_geom"] = geom
    return p


class LKA3d_deform(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.conv0 = nn.Conv3d(dim, dim, 5, padding=2, groups=dim)
        self.conv_spatial = nn.Conv3d(dim, dim, 7, stride=1, padding=9, groups=dim, dilation=3)
        self.deform_conv = DeformConvPack(in_channels=dim, out_channels=dim, kernel_size=(3, 3, 3), stride=1, padding=1)
        self.conv1 = nn.Conv3d(dim, dim, 1)

    def forward(self, x):
        return ops.lka3d_deform_forward(_block3d_params(self), x)


class LKA_Attention3d_deform(nn.Module):
    def __init__(self, d_model):
        super().__init__()
        self.proj_1 = nn.Conv3d(d_model, d_model, 1)
        self.activation = nn.GELU()
        self.spatial_gating_unit = LKA3d_deform(d_model)
        self.proj_2 = nn.Conv3d(d_model, d_model, 1)

    def forward(self, x, B, C, H, W, D):
        return ops.lka_attention3d_deform_forward(_block3d_params(self.spatial_gating_unit, self), x, B, C, H, W, D)

    def host_pipe(self, depth: int = 2):
        """A streaming pipeline for `submit_host` (keeps `depth` steps in flight; see ops.HostPipe)."""
        return ops.HostPipe(self.proj_1.weight.device, depth)

    def submit_host(self, pipe, x_host, y_host, B, C, H, W, D):
        """Enqueue one forward on host tokens without waiting; call pipe.wait() (or pipe.join()) to collect."""
        pipe.submit(_block3d_params(self.spatial_gating_unit, self), x_host, y_host, B, C, H, W, D)

    def forward_host(self, x_host, B, C, H, W, D, y_host=None):
        """Same forward for HOST tokens (ideally pinned): the library pipelines per-sample H2D copies, compute and D2H
        copies over three streams.  Parameters stay on the module's CUDA device; returns a host tensor."""
        import torch
        if y_host is None:
            y_host = torch.empty_like(x_host, pin_memory=x_host.is_pinned())
        dev = self.proj_1.weight.device
        return ops.lka_attention3d_deform_forward_host(_block3d_params(self.spatial_gating_unit, self), x_host, y_host,
                                                       B, C, H, W, D, dev)

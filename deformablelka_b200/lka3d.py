"""3D drop-in modules: ``LKA3d_deform`` and ``LKA_Attention3d_deform`` with the reference's names, ctor
arguments and state_dict keys (3D/d_lka_former/network_architecture/synapse/transformerblock.py:634-673).

Inference (no gradient needed): each forward is ONE call into libdlka_b200.
Training (grad enabled and the input or a parameter requires grad): the fused call has no backward, so the forward is
composed the way the reference composes it -- stock ``nn.Conv3d`` layers around ``self.deform_conv``, whose
``DeformConvFunction`` runs the library forward and ``dlka_deform_conv3d_backward`` -- and autograd sees every parameter.
A fused module never returns a tensor that is silently cut off from the graph."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .deform_conv3d import DeformConvPack


def needs_autograd(module: nn.Module, *tensors) -> bool:
    """True when a backward pass could be asked for: grad mode on and an input or a parameter of `module` requires grad.
    The differentiable compositions are CUDA-only like everything else here (no CPU path, 3D/dcn/src/deform_conv.h:46)."""
    if not torch.is_grad_enabled():
        return False
    need = any(t is not None and t.requires_grad for t in tensors) or any(p.requires_grad for p in module.parameters())
    if need and not all(t is None or t.is_cuda for t in tensors):
        raise RuntimeError("Not implemented on the CPU (deformablelka_b200 is CUDA-only)")
    return need


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
        if needs_autograd(self, x):
            # differentiable composition (transformerblock.py:644-652); deform_conv routes to DeformConvFunction
            attn = self.conv_spatial(self.conv0(x)).contiguous()
            return x * self.conv1(self.deform_conv(attn))
        return ops.lka3d_deform_forward(_block3d_params(self), x)


class LKA_Attention3d_deform(nn.Module):
    def __init__(self, d_model):
        super().__init__()
        self.proj_1 = nn.Conv3d(d_model, d_model, 1)
        self.activation = nn.GELU()
        self.spatial_gating_unit = LKA3d_deform(d_model)
        self.proj_2 = nn.Conv3d(d_model, d_model, 1)

    def forward(self, x, B, C, H, W, D):
        if needs_autograd(self, x):
            # differentiable composition (transformerblock.py:664-673): tokens -> NCDHW view -> ... -> tokens
            v = x.permute(0, 2, 1).reshape(B, C, H, W, D)
            t = self.proj_2(self.spatial_gating_unit(self.activation(self.proj_1(v)))) + v
            return t.reshape(B, C, H * W * D).permute(0, 2, 1)
        # inference: one library call; the packed weights persist in a per-module cache and are re-packed only when a parameter changes
        return ops.lka_attention3d_deform_forward(_block3d_params(self.spatial_gating_unit, self), x, B, C, H, W, D,
                                                  cache=self.__dict__.setdefault("_dlka_pack_cache", {}))

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

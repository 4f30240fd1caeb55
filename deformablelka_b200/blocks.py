"""Enclosing transformer blocks with the reference's class names, constructor arguments and
state_dict keys.

2D  ``DWConvLKA``, ``Mlp``, ``deformableLKABlock``      2D/networks/MaxViT_deform_LKA.py:18-52,142-189
3D  ``TransformerBlock_3D_single_deform_LKA``           3D/d_lka_former/network_architecture/synapse/transformerblock.py:570-630

Both blocks run entirely inside libdlka_b200 (one call each).  The 3D block's UnetResBlock / conv8 tail (row N3 of
SURVEY.md 8f) runs on the zero-copy tcgen05 conv kernel with the inference-mode BatchNorm folded into the weights;
the sub-modules keep monai's parameter names so that checkpoints load.
Inference only: dropout / drop-path must be 0 (the reference's eval behaviour).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .deformable_LKA import _block2d_params, deformable_LKA_Attention
from .lka3d import LKA_Attention3d_deform, _block3d_params, needs_autograd


def _pixel_shuffle_tokens(x, B, H, W, p):
    """"b h w (p1 p2 c) -> b (h p1) (w p2) c" on tokens [B, H*W, p*p*c] (differentiable training path of the patch expansion)."""
    c = x.shape[-1] // (p * p)
    return x.view(B, H, W, p, p, c).permute(0, 1, 3, 2, 4, 5).reshape(B, H * p * W * p, c)


class DWConvLKA(nn.Module):
    def __init__(self, dim=768):
        super().__init__()
        self.dwconv = nn.Conv2d(dim, dim, 3, 1, 1, bias=True, groups=dim)


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0., linear=False):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        if linear:
            raise NotImplementedError("Mlp(linear=True) (extra ReLU) is not on the D-LKA Net path")
        if act_layer is not nn.GELU:
            raise NotImplementedError("only act_layer=nn.GELU (the reference default) is implemented")
        if out_features != in_features:
            raise NotImplementedError("out_features != in_features is not used by deformableLKABlock")
        self.fc1 = nn.Conv2d(in_features, hidden_features, 1)
        self.dwconv = DWConvLKA(hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Conv2d(hidden_features, out_features, 1)
        self.drop = nn.Dropout(drop)
        self.linear = linear
        self.hidden_features = hidden_features


class deformableLKABlock(nn.Module):
    def __init__(self, dim, mlp_ratio=4., drop=0., drop_path=0., act_layer=nn.GELU, linear=False):
        super().__init__()
        if drop != 0. or drop_path != 0.:
            raise NotImplementedError("forward-only build: drop / drop_path must be 0 (eval behaviour)")
        self.norm1 = nn.LayerNorm(dim)
        self.attn = deformable_LKA_Attention(dim)
        self.drop_path = nn.Identity()
        self.norm2 = nn.LayerNorm(dim)
        mlp_hidden_dim = int(dim * mlp_ratio)
        self.mlp = Mlp(in_features=dim, hidden_features=mlp_hidden_dim, act_layer=act_layer, drop=drop, linear=linear)
        layer_scale_init_value = 1e-2
        self.layer_scale_1 = nn.Parameter(layer_scale_init_value * torch.ones((dim)), requires_grad=True)
        self.layer_scale_2 = nn.Parameter(layer_scale_init_value * torch.ones((dim)), requires_grad=True)

    def forward_autograd(self, x, H, W):
        """Training / gradient path: MaxViT_deform_LKA.py:165-189 composed from this module's own sub-modules."""
        B, N, C = x.shape
        v = x.permute(0, 2, 1).reshape(B, C, H, W)
        y = self.attn(self.norm1(v.permute(0, 2, 3, 1)).permute(0, 3, 1, 2))
        v = v + self.layer_scale_1.unsqueeze(-1).unsqueeze(-1) * self.drop_path(y)
        y = self.norm2(v.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)
        m = self.mlp
        y = m.fc2(m.drop(m.act(m.dwconv.dwconv(m.fc1(y)))))
        v = v + self.layer_scale_2.unsqueeze(-1).unsqueeze(-1) * self.drop_path(m.drop(y))
        return v.reshape(B, C, N).permute(0, 2, 1)

    def forward(self, x, H, W):
        if needs_autograd(self, x):
            return self.forward_autograd(x, H, W)
        blk = {
            "norm1_weight": self.norm1.weight, "norm1_bias": self.norm1.bias, "layer_scale_1": self.layer_scale_1,
            "norm2_weight": self.norm2.weight, "norm2_bias": self.norm2.bias, "layer_scale_2": self.layer_scale_2,
            "fc1_weight": self.mlp.fc1.weight, "fc1_bias": self.mlp.fc1.bias,
            "dw_weight": self.mlp.dwconv.dwconv.weight, "dw_bias": self.mlp.dwconv.dwconv.bias,
            "fc2_weight": self.mlp.fc2.weight, "fc2_bias": self.mlp.fc2.bias,
        }
        return ops.deformable_lka_block2d_forward(_block2d_params(self.attn.spatial_gating_unit, self.attn), blk, x, H, W,
                                                  self.mlp.hidden_features, self.norm1.eps, self.norm2.eps)


class PatchExpand(nn.Module):
    """2D/networks/MaxViT_deform_LKA.py:488-513 -- Linear(dim, 2*dim, bias=False) -> pixel shuffle x2 -> LayerNorm(dim/2)."""

    def __init__(self, input_resolution, dim, dim_scale=2, norm_layer=nn.LayerNorm):
        super().__init__()
        if dim_scale != 2:
            raise NotImplementedError("PatchExpand: only dim_scale=2 (the only value the reference constructs, :574)")
        self.input_resolution = input_resolution
        self.dim = dim
        self.expand = nn.Linear(dim, 2 * dim, bias=False)
        self.norm = norm_layer(dim // dim_scale)

    def forward(self, x):
        H, W = self.input_resolution
        if needs_autograd(self, x):   # training: stock layers, autograd sees expand / norm
            assert x.shape[1] == H * W, "input feature has wrong size"
            return self.norm(_pixel_shuffle_tokens(self.expand(x), x.shape[0], H, W, 2))
        return ops.patch_expand2d_forward(x, self.expand.weight, self.norm.weight, self.norm.bias, self.norm.eps, H, W, 2)


class FinalPatchExpand_X4(nn.Module):
    """:516-545 -- Linear(dim, 16*dim, bias=False) -> pixel shuffle x4 -> LayerNorm(dim)."""

    def __init__(self, input_resolution, dim, dim_scale=4, norm_layer=nn.LayerNorm):
        super().__init__()
        if dim_scale != 4:
            raise NotImplementedError("FinalPatchExpand_X4: dim_scale must be 4")
        self.input_resolution = input_resolution
        self.dim = dim
        self.dim_scale = dim_scale
        self.expand = nn.Linear(dim, 16 * dim, bias=False)
        self.output_dim = dim
        self.norm = norm_layer(self.output_dim)

    def forward(self, x):
        H, W = self.input_resolution
        if needs_autograd(self, x):
            assert x.shape[1] == H * W, "input feature has wrong size"
            return self.norm(_pixel_shuffle_tokens(self.expand(x), x.shape[0], H, W, 4))
        return ops.patch_expand2d_forward(x, self.expand.weight, self.norm.weight, self.norm.bias, self.norm.eps, H, W, 4)


class MyDecoderLayer(nn.Module):
    """One 2D decoder stage (:548-620): x1_linear + skip add, two deformableLKABlocks, patch expansion, and on the last stage
    the 1x1 class head.  Same constructor arguments and parameter names as the reference; five library calls."""

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
        for m in self.modules():   # init_weights, :585-598
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.LayerNorm):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Conv2d):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def forward(self, x1, x2=None):
        if x2 is None:
            return self.layer_up(x1)
        b, h, w, c = x2.shape
        grad = needs_autograd(self, x1, x2)
        if grad:   # every sub-module takes its differentiable path
            cat_linear_x = self.x1_linear(x1) + x2.reshape(b, -1, c)
        else:
            cat_linear_x = ops.linear_tokens_forward(x1, self.x1_linear.weight, self.x1_linear.bias, add=x2.reshape(b, -1, c))
        t = self.layer_lka_2(self.layer_lka_1(cat_linear_x, h, w), h, w)
        if self.last_layer is None:
            return self.layer_up(t)
        up = self.layer_up(t)                                     # [b, 16*h*w, out_dim] tokens = NHWC
        if grad:
            return self.last_layer(up.view(b, 4 * h, 4 * w, -1).permute(0, 3, 1, 2))
        logits = ops.linear_tokens_forward(up, self.last_layer.weight.flatten(1), self.last_layer.bias)
        return logits.view(b, 4 * h, 4 * w, -1).permute(0, 3, 1, 2).contiguous()   # NCHW like nn.Conv2d's output


class _ConvNoBias3d(nn.Module):
    """monai ``Convolution`` keeps its conv under ``.conv`` -> parameter key ``<name>.conv.weight``."""

    def __init__(self, cin, cout, k):
        super().__init__()
        self.conv = nn.Conv3d(cin, cout, k, stride=1, padding=k // 2, bias=False)

    def forward(self, x):
        return self.conv(x)


class UnetResBlock(nn.Module):
    """Stock-PyTorch stand-in for monai's UnetResBlock(3, C, C, kernel_size=3, stride=1, norm_name="batch")
    (3D/d_lka_former/network_architecture/dynunet_block.py:12-80): same parameter names.  Row N3: not native yet."""

    def __init__(self, spatial_dims, in_channels, out_channels, kernel_size=3, stride=1, norm_name="batch"):
        super().__init__()
        assert spatial_dims == 3 and stride == 1 and norm_name == "batch" and in_channels == out_channels
        self.conv1 = _ConvNoBias3d(in_channels, out_channels, kernel_size)
        self.conv2 = _ConvNoBias3d(out_channels, out_channels, kernel_size)
        self.lrelu = nn.LeakyReLU(negative_slope=0.01, inplace=True)
        self.norm1 = nn.BatchNorm3d(out_channels)
        self.norm2 = nn.BatchNorm3d(out_channels)

    def forward(self, inp):
        out = self.lrelu(self.norm1(self.conv1(inp)))
        out = self.norm2(self.conv2(out))
        return self.lrelu(out + inp)


class TransformerBlock_3D_single_deform_LKA(nn.Module):
    def __init__(self, input_size: int, hidden_size: int, proj_size: int, num_heads: int, dropout_rate: float = 0.0,
                 pos_embed=False) -> None:
        super().__init__()
        if not (0 <= dropout_rate <= 1):
            raise ValueError("dropout_rate should be between 0 and 1.")
        if hidden_size % num_heads != 0:
            raise ValueError("hidden_size should be divisible by num_heads.")
        self.norm = nn.LayerNorm(hidden_size)
        self.gamma = nn.Parameter(1e-6 * torch.ones(hidden_size), requires_grad=True)
        self.epa_block = self._attention_class()(d_model=hidden_size)
        self.conv51 = UnetResBlock(3, hidden_size, hidden_size, kernel_size=3, stride=1, norm_name="batch")
        self.conv8 = nn.Sequential(nn.Dropout3d(0.1, False), nn.Conv3d(hidden_size, hidden_size, 1))
        self.pos_embed = None
        if pos_embed:
            self.pos_embed = nn.Parameter(torch.zeros(1, input_size, hidden_size))

    @staticmethod
    def _attention_class():
        return LKA_Attention3d_deform   # the ACDC network's block (acdc.py) swaps in its own stencil shapes

    def attention_half(self, x_tokens, B, C, H, W, D):
        """x' = x + pos_embed; x' + gamma * epa_block(norm(x'))  -- one library call (transformerblock.py:620-624)."""
        if needs_autograd(self, x_tokens):
            raise RuntimeError("attention_half is a fused inference entry and a gradient is required: use forward()")
        ep = self.epa_block
        return ops.lka_transformer3d_prenorm_forward(_block3d_params(ep.spatial_gating_unit, ep), self.norm.weight,
                                                     self.norm.bias, self.norm.eps, self.gamma, self.pos_embed, x_tokens,
                                                     B, C, H, W, D)

    @staticmethod
    def _fold_bn(bn: nn.BatchNorm3d):
        """Inference-mode BatchNorm as per-channel (scale, shift) -- a few [C]-sized host-side tensor ops."""
        scale = bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)
        return scale.contiguous(), (bn.bias.detach() - bn.running_mean * scale).contiguous()

    def forward_tokens(self, tokens, B, C, H, W, D):
        """The whole block on tokens [B, N, C] in ONE library call (rows N1 + N3): attention half, UnetResBlock
        (two 3x3x3 convs on tcgen05 with folded BatchNorm + LeakyReLU + residual) and conv8 + residual."""
        if self.training:
            raise RuntimeError("forward_tokens is the fused inference entry: call .eval() (BatchNorm / Dropout3d), or use "
                               "forward(), which takes the differentiable path in training mode")
        if needs_autograd(self, tokens):
            raise RuntimeError("forward_tokens is the fused inference entry and a gradient is required: wrap it in "
                               "torch.no_grad(), or use forward(), which takes the differentiable path")
        ep = self.epa_block
        s1, t1 = self._fold_bn(self.conv51.norm1)
        s2, t2 = self._fold_bn(self.conv51.norm2)
        tail = {"norm_weight": self.norm.weight, "norm_bias": self.norm.bias, "gamma": self.gamma, "pos_embed": self.pos_embed,
                "conv1_weight": self.conv51.conv1.conv.weight, "bn1_scale": s1, "bn1_shift": t1,
                "conv2_weight": self.conv51.conv2.conv.weight, "bn2_scale": s2, "bn2_shift": t2,
                "conv8_weight": self.conv8[1].weight, "conv8_bias": self.conv8[1].bias}
        return ops.lka_transformer3d_block_forward(_block3d_params(ep.spatial_gating_unit, ep), tail, self.norm.eps,
                                                   self.conv51.lrelu.negative_slope, tokens, B, C, H, W, D)

    def forward(self, x):
        B, C, H, W, D = x.shape
        if self.training or needs_autograd(self, x):
            return self.forward_autograd(x)
        tokens = x.reshape(B, C, H * W * D).permute(0, 2, 1).contiguous()          # (B, N, C): layout plumbing only
        out = self.forward_tokens(tokens, B, C, H, W, D)
        return out.reshape(B, H, W, D, C).permute(0, 4, 1, 2, 3)                    # (B, C, H, W, D) view, as the reference

    def forward_autograd(self, x):
        """Training / gradient path (transformerblock.py:617-630 composed from this module's own sub-modules): LayerNorm, the
        attention block's differentiable composition (DeformConvFunction inside), BatchNorm / Dropout3d in their current mode."""
        B, C, H, W, D = x.shape
        t = x.reshape(B, C, H * W * D).permute(0, 2, 1)
        if self.pos_embed is not None:
            t = t + self.pos_embed
        attn = t + self.gamma * self.epa_block(self.norm(t), B, C, H, W, D)
        attn_skip = attn.reshape(B, H, W, D, C).permute(0, 4, 1, 2, 3)
        return attn_skip + self.conv8(self.conv51(attn_skip))

    def forward_reference_tail(self, x):
        """Attention half native, UnetResBlock / conv8 through stock PyTorch layers (cross-check of row N3)."""
        B, C, H, W, D = x.shape
        tokens = x.reshape(B, C, H * W * D).permute(0, 2, 1).contiguous()
        attn = self.attention_half(tokens, B, C, H, W, D)
        attn_skip = attn.reshape(B, H, W, D, C).permute(0, 4, 1, 2, 3)
        return attn_skip + self.conv8(self.conv51(attn_skip))

"""Small forward passes of every entry point, meant to run under compute-sanitizer (memcheck / racecheck)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import deformablelka_b200 as dl

dev = "cuda:0"
torch.manual_seed(0)
for math in os.environ.get("DLKA_SMOKE_MATH", "bf16x3,fp32").split(","):
    os.environ["DLKA_MATH"] = math
    with torch.no_grad():
        m3 = dl.LKA_Attention3d_deform(32)
        m3.spatial_gating_unit.deform_conv.conv_offset.weight.normal_(0, 0.05)
        m3.spatial_gating_unit.deform_conv.conv_offset.bias.uniform_(-2, 2)
        m3 = m3.to(dev)
        y = m3(torch.randn(1, 5 * 7 * 9, 32, device=dev), 1, 32, 5, 7, 9)          # ragged extents (tile edges)
        y = m3.spatial_gating_unit(torch.randn(1, 32, 5, 7, 9, device=dev))
        m2 = dl.deformable_LKA_Attention(16).to(dev)
        y = m2(torch.randn(1, 16, 9, 13, device=dev))
        b2 = dl.deformableLKABlock(16).to(dev)
        y = b2(torch.randn(1, 9 * 13, 16, device=dev), 9, 13)
        t3 = dl.TransformerBlock_3D_single_deform_LKA(5 * 7 * 9, 32, 32, 4, pos_embed=True).to(dev).eval()
        y = t3.attention_half(torch.randn(1, 5 * 7 * 9, 32, device=dev), 1, 32, 5, 7, 9)
        x = torch.randn(1, 32, 5, 6, 7, device=dev)
        y = dl.ops.deform_conv3d_forward(x, torch.randn(32, 32, 3, 3, 3, device=dev), torch.randn(32, device=dev),
                                         torch.randn(1, 81, 5, 6, 7, device=dev) * 4, 3, 1, 1, 1, 1, 1)
        xh = torch.randn(2, 5 * 7 * 9, 32).pin_memory()
        y = m3.forward_host(xh, 2, 32, 5, 7, 9)
        y = t3(torch.randn(1, 32, 5, 7, 9, device=dev)) if hasattr(t3, "forward") else None        # whole block incl. UnetResBlock (N3)
        for dim, dims in ((32, (4, 9, 7)), (128, (3, 6, 5)), (256, (2, 4, 3))):                       # ACDC stencil shapes (N4)
            ma = dl.acdc.LKA_Attention3d_deform(dim)
            ma.spatial_gating_unit.deform_conv.conv_offset.bias.uniform_(-2, 2)
            ma = ma.to(dev)
            y = ma(torch.randn(1, dims[0] * dims[1] * dims[2], dim, device=dev), 1, dim, *dims)
        dec = dl.MyDecoderLayer((5, 6), [32] * 5, 1, "mix_skip", n_class=9, is_last=True).to(dev)   # 2D decoder stage (N3)
        y = dec(torch.randn(1, 30, 32, device=dev), torch.randn(1, 5, 6, 32, device=dev))
        g = dl.ops.deform_conv3d_backward(x, torch.randn(32, 32, 3, 3, 3, device=dev), torch.randn(32, device=dev),   # N2
                                          torch.randn(1, 81, 5, 6, 7, device=dev) * 4, torch.randn(1, 32, 5, 6, 7, device=dev),
                                          3, 1, 1, 1, 1, 1)
        # round-2 kernels: K-split offset conv + thread-per-voxel stencil (small grid, C = 64), persistent 1x1 streaming kernel
        # (>= 2 tiles per SM), grouped 3D backward, 2D operator backward
        m64 = dl.LKA_Attention3d_deform(64)
        m64.spatial_gating_unit.deform_conv.conv_offset.bias.uniform_(-2, 2)
        m64 = m64.to(dev)
        y = m64(torch.randn(1, 8 * 8 * 8, 64, device=dev), 1, 64, 8, 8, 8)
        y = dl.ops.linear_tokens_forward(torch.randn(40000, 32, device=dev), torch.randn(32, 32, device=dev), torch.randn(32, device=dev))
        xg = torch.randn(1, 16, 4, 5, 6, device=dev)
        g = dl.ops.deform_conv3d_backward(xg, torch.randn(16, 8, 3, 3, 3, device=dev), torch.randn(16, device=dev),
                                          torch.randn(1, 2 * 81, 4, 5, 6, device=dev), torch.randn(1, 16, 4, 5, 6, device=dev),
                                          3, 1, 1, 1, 2, 2)
        x2 = torch.randn(1, 16, 9, 11, device=dev)
        g = dl.ops.deform_conv2d_backward(x2, torch.randn(1, 50, 9, 11, device=dev), torch.randn(16, 1, 5, 5, device=dev), None,
                                          torch.randn(1, 16, 9, 11, device=dev), padding=2)
    torch.cuda.synchronize()
    print("ok", math, flush=True)

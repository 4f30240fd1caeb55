"""Small forward passes of every entry point, meant to run under compute-sanitizer (memcheck / racecheck)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import deformablelka_b200 as dl

dev = "cuda:0"
torch.manual_seed(0)
for math in ("bf16x3", "fp32"):
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
    torch.cuda.synchronize()
    print("ok", math, flush=True)

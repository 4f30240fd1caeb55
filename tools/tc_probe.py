"""GPU probe: tensor-core (bf16x3) path vs the exact fp32 SIMT path, op by op.  Run under `timeout`."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import deformablelka_b200 as dl

dev = "cuda:0"


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max()).item()


def probe_deform(C, Co, dims, scale=1.0):
    torch.manual_seed(0)
    B, (D, H, W) = 2, dims
    x = torch.randn(B, C, D, H, W, device=dev); w = torch.randn(Co, C, 3, 3, 3, device=dev) * 0.1
    b = torch.randn(Co, device=dev); off = torch.randn(B, 81, D, H, W, device=dev) * scale
    ref = dl.ops.deform_conv3d_forward(x, w, b, off, 3, 1, 1, 1, 1, 1, math="fp32")
    got = dl.ops.deform_conv3d_forward(x, w, b, off, 3, 1, 1, 1, 1, 1, math="bf16x3")
    torch.cuda.synchronize()
    print(f"deform op   C={C:3d} Co={Co:3d} dims={dims} scale={scale}: rel {rel(got, ref):.2e}", flush=True)


def probe_pack(C, dims):
    torch.manual_seed(1)
    B, (D, H, W) = 2, dims
    x = torch.randn(B, C, D, H, W, device=dev)
    ow = torch.randn(81, C, 3, 3, 3, device=dev) * 0.05; ob = torch.rand(81, device=dev) * 2 - 1
    w = torch.randn(C, C, 3, 3, 3, device=dev) * 0.1; b = torch.randn(C, device=dev)
    ref = dl.ops.deform_conv_pack3d(x, ow, ob, w, b, 1, 1, 1, 1, 1, math="fp32")
    got = dl.ops.deform_conv_pack3d(x, ow, ob, w, b, 1, 1, 1, 1, 1, math="bf16x3")
    torch.cuda.synchronize()
    print(f"pack op     C={C:3d} dims={dims}: rel {rel(got, ref):.2e}", flush=True)


def probe_block3d(C, dims):
    from oracle import oracle
    torch.manual_seed(2)
    m = dl.LKA_Attention3d_deform(C)
    oracle.randomize_offsets_(m)
    m = m.to(dev)
    H, W, D = dims
    x = torch.randn(2, H * W * D, C, device=dev)
    os.environ["DLKA_MATH"] = "fp32"; ref = m(x, 2, C, H, W, D)
    os.environ["DLKA_MATH"] = "bf16x3"; got = m(x, 2, C, H, W, D)
    torch.cuda.synchronize()
    print(f"block3d     C={C:3d} dims={dims}: rel {rel(got, ref):.2e}", flush=True)


def probe_block2d(C, hw):
    torch.manual_seed(3)
    m = dl.deformable_LKA_Attention(C).to(dev)
    x = torch.randn(2, C, *hw, device=dev)
    os.environ["DLKA_MATH"] = "fp32"; ref = m(x)
    os.environ["DLKA_MATH"] = "bf16x3"; got = m(x)
    torch.cuda.synchronize()
    print(f"block2d     C={C:3d} hw={hw}: rel {rel(got, ref):.2e}", flush=True)


if __name__ == "__main__":
    probe_deform(32, 32, (4, 8, 8))
    probe_deform(96, 96, (6, 7, 9), 3.0)
    probe_deform(64, 48, (5, 5, 5))
    probe_deform(128, 128, (4, 4, 4))
    probe_deform(256, 256, (4, 4, 4))
    probe_pack(32, (8, 8, 8))
    probe_pack(96, (6, 10, 7))
    for C in (32, 64, 96, 128, 256):
        probe_block3d(C, (8, 6, 10) if C < 128 else (4, 4, 4))
    for C in (64, 96, 192, 384):
        probe_block2d(C, (14, 14))
    # large-M path (MT=2): 2*128*148*2 rows or more
    probe_block3d(32, (48, 48, 40))
    t0 = time.time(); probe_block3d(96, (32, 64, 64)); print("  took", time.time() - t0)
    print("PROBE DONE")

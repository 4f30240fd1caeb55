"""Times the non-headline BASELINE.json configurations on one GPU (CUDA events, 3 warm-ups, inputs resident):
C2  2D D-LKA Net block shapes (batch 24):   [24,384,14,14] [24,192,28,28] [24,96,56,56]  (x2 blocks each in the net)
C3  3D deformable conv op alone (2,64,32,64,64), k=3, explicit offsets
C4  3D D-LKA Net block shapes (batch 2):    (32,32^3) (64,16^3) (128,8^3) (256,4^3)
Prints one JSON line per case.  Not part of the driver contract (bench.py is)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import deformablelka_b200 as dl

dev = "cuda:0"


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def breakdown(fn, iters=10):
    """per-kernel device time (us per call) from the library's event profiler"""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dl._lib.profile_enable(True)
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    prof = dl._lib.profile_summary()
    dl._lib.profile_enable(False)
    return {k: round(v[1] / iters * 1e3, 1) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}


def main():
    math = os.environ.get("DLKA_MATH", "bf16x3")
    os.environ["DLKA_MATH"] = math
    torch.manual_seed(1234)
    out = []
    with torch.no_grad():
        for C, hw in ((384, 14), (192, 28), (96, 56), (64, 224)):
            B = 24 if hw != 224 else 1
            m = dl.deformable_LKA_Attention(C).to(dev).eval()
            x = torch.randn(B, C, hw, hw, device=dev)
            ms = timeit(lambda: m(x))
            out.append({"config": "C2" if hw != 224 else "C1-shape", "op": "deformable_LKA_Attention", "shape": [B, C, hw, hw], "ms": ms,
                        "Mpx_per_s": B * hw * hw / ms / 1e3, "math": math, "kernels_us": breakdown(lambda: m(x))})
        B, C, D, H, W = 2, 64, 32, 64, 64
        x = torch.randn(B, C, D, H, W, device=dev); w = torch.randn(C, C, 3, 3, 3, device=dev) * 0.05
        b = torch.randn(C, device=dev); off = torch.randn(B, 81, D, H, W, device=dev)
        ms = timeit(lambda: dl.ops.deform_conv3d_forward(x, w, b, off, 3, 1, 1, 1, 1, 1, 64))
        out.append({"config": "C3", "op": "deform_conv3d_forward (NCDHW in/out)", "shape": [B, C, D, H, W], "ms": ms,
                    "GVoxel_per_s": B * D * H * W / ms / 1e6, "math": math})
        for C, s in ((32, 32), (64, 16), (128, 8), (256, 4)):
            m = dl.LKA_Attention3d_deform(C)
            co = m.spatial_gating_unit.deform_conv.conv_offset
            co.weight.normal_(0, 0.05); co.bias.uniform_(-1, 1)
            m = m.to(dev).eval()
            x = torch.randn(2, s * s * s, C, device=dev)
            ms = timeit(lambda: m(x, 2, C, s, s, s))
            out.append({"config": "C4", "op": "LKA_Attention3d_deform", "shape": [2, C, s, s, s], "ms": ms,
                        "GVoxel_per_s": 2 * s ** 3 / ms / 1e6, "math": math, "kernels_us": breakdown(lambda: m(x, 2, C, s, s, s))})
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()

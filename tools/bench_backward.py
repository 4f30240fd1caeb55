"""Row N2 measurement: forward + backward of the 3D deformable conv operator at BASELINE config 3 (2,64,32,64,64), k = 3,
through the public Python API (DeformConvFunction), CUDA events, beside the reference's compiled D3D when oracle/_ref exists.
    python tools/bench_backward.py [--steps 5]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import deformablelka_b200 as dl

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--shape", type=int, nargs=5, default=[2, 64, 32, 64, 64])
args = ap.parse_args()
dev = torch.device("cuda", 0)
B, C, D, H, W = args.shape
torch.manual_seed(0)
x = torch.randn(B, C, D, H, W, device=dev)
w = torch.randn(C, C, 3, 3, 3, device=dev) * 0.05
b = torch.randn(C, device=dev)
off = torch.randn(B, 81, D, H, W, device=dev) * 0.5
gout = torch.randn(B, C, D, H, W, device=dev)


def timed(fn, steps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


out = {"shape": args.shape, "voxels": B * D * H * W}
out["forward_ms"] = timed(lambda: dl.ops.deform_conv3d_forward(x, w, b, off, 3, 1, 1, 1, 1, 1, 64), args.steps)
out["backward_ms"] = timed(lambda: dl.ops.deform_conv3d_backward(x, w, b, off, gout, 3, 1, 1, 1, 1, 1, 64), args.steps)
try:
    from oracle import build_ref
    d3d = build_ref.load_d3d()
except Exception:
    d3d = None
if d3d is not None:
    torch.backends.cuda.matmul.allow_tf32 = False
    g = (3, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 64)
    out["reference_forward_ms"] = timed(lambda: d3d.deform_conv_forward(x, w, b, off, *g), args.steps)
    out["reference_backward_ms"] = timed(lambda: d3d.deform_conv_backward(x, w, b, off, gout, *g), args.steps)
print(json.dumps(out))
# per-kernel breakdown of one backward call (CUDA events inside the library)
dl._lib.profile_enable(True)
dl.ops.deform_conv3d_backward(x, w, b, off, gout, 3, 1, 1, 1, 1, 1, 64)
torch.cuda.synchronize()
prof = dl._lib.profile_summary()
dl._lib.profile_enable(False)
print(json.dumps({k: [v[0], round(v[1], 3)] for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}))

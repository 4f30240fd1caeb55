"""Debug timeline of one CTA of the deformable-conv kernel (clock64 stamps written by the kernel itself).
Prints per-K-step intervals of every pipeline role, to see which hand-off paces the loop."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes

import torch

import deformablelka_b200 as dl
from bench import make_block

dev = torch.device("cuda", 0)
os.environ["DLKA_MATH"] = "bf16x3"
B, C, D1, D2, D3 = 2, 96, 64, 128, 128
m = make_block(C, dev)
x = torch.randn(B, D1 * D2 * D3, C, device=dev)
buf = torch.zeros(6, 1024, 2, dtype=torch.int64, device=dev)
L = dl._lib.lib
L.dlka_debug_deform_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
with torch.no_grad():
    for _ in range(2):
        m(x, B, C, D1, D2, D3)
    torch.cuda.synchronize()
    L.dlka_debug_deform_trace(buf.data_ptr(), int(sys.argv[1]) if len(sys.argv) > 1 else 5000)
    m(x, B, C, D1, D2, D3)
    torch.cuda.synchronize()
    L.dlka_debug_deform_trace(None, 0)
t = buf.cpu()
KS = 81
t0 = int(t[0, 0, 0])
names = ["mma B-ready", "mma A-ready", "ldr slot-free", "prm start", "prm done", "g8 waits-done", "g8 arrive", "g23 waits-done",
         "g23 arrive", "g8 loads-used", "g8 fullP-ok"]
cols = [(0, 0), (0, 1), (1, 0), (2, 0), (2, 1), (3, 0), (3, 1), (4, 0), (4, 1), (5, 0), (5, 1)]
print("ks " + " ".join(f"{n:>14s}" for n in names))
for ks in list(range(0, 12)) + list(range(40, 46)) + list(range(76, 81)):
    print(f"{ks:2d} " + " ".join(f"{int(t[r, ks, e]) - t0:14d}" for r, e in cols))
a_ready = t[0, :KS, 1]
print("per-ks period (mma A-ready deltas): mean", float((a_ready[1:] - a_ready[:-1]).float().mean()),
      " total main loop", int(a_ready[KS - 1] - a_ready[0]))
g = t[3, :KS]
print("gather warp 8: mean busy (waits-done -> arrive)", float((g[:, 1] - g[:, 0]).float().mean()),
      " mean wait (arrive -> next waits-done)", float((g[1:, 0] - g[:-1, 1]).float().mean()))
p = t[2, :KS]
print("param warp 4: mean busy", float((p[:, 1] - p[:, 0]).float().mean()), " mean wait", float((p[1:, 0] - p[:-1, 1]).float().mean()))
mm = t[0, :KS]
print("mma: mean wait for A after B ready", float((mm[:, 1] - mm[:, 0]).float().mean()))
lu = t[5, :KS]
print("g8: waits-done -> loads consumed", float((lu[:, 0] - g[:, 0]).float().mean()), "; loads consumed -> arrive", float((g[:, 1] - lu[:, 0]).float().mean()))
print("g8: fullP-ok -> emptyA ok", float((g[:, 0] - lu[:, 1]).float().mean()))

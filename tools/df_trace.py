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
import numpy as np


def stats(role, label):
    """busy = event 0 -> event 1 of the same K step; gap = event 1 -> event 0 of the warp's next K step (gather warps
    take every DF_GROUPS-th step, so untouched entries are zero and skipped)."""
    r = t[role, :KS].numpy()
    ks = np.nonzero(r[:, 0])[0]
    busy = (r[ks, 1] - r[ks, 0]).mean()
    gap = (r[ks[1:], 0] - r[ks[:-1], 1]).mean() if len(ks) > 1 else float("nan")
    print(f"{label}: steps {len(ks)}  mean busy {busy:.1f}  mean wait {gap:.1f}")
    return r, ks


g, gks = stats(3, "gather warp 8  (waits-done -> arrive)")
stats(4, "gather warp 23 (waits-done -> arrive)")
stats(2, "param warp 4")
mm = t[0, :KS].numpy()
print("mma: mean wait for A after B ready", (mm[:, 1] - mm[:, 0]).mean())
lu = t[5, :KS].numpy()
print("g8: waits-done -> loads consumed", (lu[gks, 0] - g[gks, 0]).mean(), "; loads consumed -> arrive", (g[gks, 1] - lu[gks, 0]).mean())
print("g8: fullP-ok -> emptyA ok", (g[gks, 0] - lu[gks, 1]).mean())
misc = t[1, :32, 1].numpy()
lab = {0: "entry", 1: "setup done", 2: "accFull seen", 3: "stage0 start", 4: "stage0 done", 5: "stage1 start (conv1 acc ready)",
       6: "stage1 done", 7: "stage2 start (proj_2 acc ready)", 8: "stage2 done (stored)", 10: "exit barrier passed"}
print("CTA timeline (cycles from kernel entry of this CTA; first mma B-ready at", t0 - int(misc[0]), ")")
for k in sorted(lab):
    if misc[k]:
        print(f"  {lab[k]:34s} {int(misc[k] - misc[0]):8d}")
print("  last main-loop A-ready              ", int(t[0, KS - 1, 1] - misc[0]))
for st in (1, 2):
    print(f"  mma stage {st}: weights ready {int(misc[20 + 3 * st] - misc[0])}  A operand ready {int(misc[21 + 3 * st] - misc[0])}  "
          f"MMAs issued {int(misc[22 + 3 * st] - misc[0])}")

"""Why is the e2e step (18.8 ms) longer than both the device step (17.1 ms) and the two PCIe copies run alone (16.1 ms)?
Runs the device-resident step with and without the same copies in flight on two side streams and reports both sides."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
import deformablelka_b200 as dl

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
os.environ["DLKA_MATH"] = "bf16x3"
dl.ops.bind_host_thread(dev)
B, C, D1, D2, D3 = 2, 96, 64, 128, 128
N = D1 * D2 * D3
m = bench.make_block(C, dev)
x = torch.randn(B, N, C, device=dev)
n = B * N * C
hx = dl.ops.pinned_empty((n,), dev); hy = dl.ops.pinned_empty((n,), dev)
hx.zero_(); hy.zero_()
dx = torch.empty(n, device=dev); dy = torch.zeros(n, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
K = 8


def run(copies: bool):
    with torch.no_grad():
        for _ in range(3):
            m(x, B, C, D1, D2, D3)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    ev[0].record(cur)
    if copies:
        ev[2].record(s1); ev[4].record(s2)
        for _ in range(K):
            with torch.cuda.stream(s1):
                dx.copy_(hx, non_blocking=True)
            with torch.cuda.stream(s2):
                hy.copy_(dy, non_blocking=True)
        ev[3].record(s1); ev[5].record(s2)
    with torch.no_grad():
        for _ in range(K):
            m(x, B, C, D1, D2, D3)
    ev[1].record(cur)
    torch.cuda.synchronize()
    out = {"compute_ms_per_step": ev[0].elapsed_time(ev[1]) / K}
    if copies:
        out["h2d_ms_per_step"] = ev[2].elapsed_time(ev[3]) / K
        out["d2h_ms_per_step"] = ev[4].elapsed_time(ev[5]) / K
    return out


print(json.dumps({"alone": run(False), "with_both_copies_in_flight": run(True)}))

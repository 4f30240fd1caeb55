"""Ceiling of the e2e leg: one step moves 805 MB host->device and 805 MB device->host.  Times those two copies alone
(pinned, NUMA-placed like bench.py's buffers), one direction at a time and both directions at once on two streams."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import deformablelka_b200 as dl

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
node = dl.ops.bind_host_thread(dev)
n = 2 * 64 * 128 * 128 * 96
hx = dl.ops.pinned_empty((n,), dev)
hy = dl.ops.pinned_empty((n,), dev)
hx.zero_(); hy.zero_()
dx = torch.empty(n, device=dev); dy = torch.zeros(n, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, iters=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def h2d():
    s1.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        dx.copy_(hx, non_blocking=True)


def d2h():
    s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s2):
        hy.copy_(dy, non_blocking=True)


def both():
    h2d(); d2h()


gb = n * 4 / 1e9
out = {"bytes_per_direction": n * 4, "numa_node": node}
for name, fn in (("h2d_only", h2d), ("d2h_only", d2h), ("both_directions", both)):
    ms = timed(fn)
    out[name] = {"ms": ms, "GB_per_s_per_direction": gb / (ms * 1e-3)}
print(json.dumps(out))

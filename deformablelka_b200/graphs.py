"""CUDA-graph replay of an inference call (small shapes: the deep stages of both networks).

At 4^3 .. 16^3 volumes and 14^2 images a block is 5-8 kernels of 10-50 us each, so the host side of a call (Python module ->
ctypes -> planning -> launches) is as long as the device work.  ``GraphedCall`` captures ONE call of any library-backed module
on static buffers and replays it with a single ``cudaGraphLaunch``: the library allocates nothing on the device, launches only
on the caller's stream and keeps its workspace / packed weights in buffers that outlive the call, so its calls are capturable
as they are (the reference has no counterpart: it relies on eager PyTorch launches).

    g = GraphedCall(block, x_example, B, C, H, W, D)      # warm-up (workspace, packed weights), then capture
    y = g(x)                                               # copy-in, replay; returns the static output tensor

Inference only (the fused entries are; torch.no_grad is entered for capture and replay).  A parameter update after capture is
NOT seen by the graph when the packed-weight cache is in use -- call ``recapture()`` after loading new weights.
"""
from __future__ import annotations

import torch

from ._lib import Workspace


class GraphedCall:
    def __init__(self, fn, *example_args, warmup: int = 3, **kwargs):
        tensors = [a for a in example_args if torch.is_tensor(a)]
        if not tensors or not all(t.is_cuda for t in tensors):
            raise RuntimeError("Not implemented on the CPU (GraphedCall captures a CUDA graph)")
        self.fn, self.kwargs, self.warmup = fn, kwargs, warmup
        self.static_args = [a.clone() if torch.is_tensor(a) else a for a in example_args]
        self.graph = None
        self.static_out = None
        self.recapture()

    def recapture(self):
        """(Re-)record the graph: needed after the parameters of the captured module changed."""
        dev = next(a for a in self.static_args if torch.is_tensor(a)).device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(self.warmup):   # allocates the cached workspace and packs the weights OUTSIDE the capture
                self.fn(*self.static_args, **self.kwargs)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        # the library's scratch buffer is cached per (device, stream): the warm-up stream's copy is not needed again ...
        Workspace._bufs.pop((dev.index if dev.index is not None else torch.cuda.current_device(), side.cuda_stream), None)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_out = self.fn(*self.static_args, **self.kwargs)
            key = (dev.index if dev.index is not None else torch.cuda.current_device(), torch.cuda.current_stream(dev).cuda_stream)
        # ... and the one allocated during capture (from the graph's private pool) is what every replay writes to: this object keeps
        # it alive, and the cache forgets it so that an unrelated call on a recycled stream cannot regrow / replace it
        self._workspace = Workspace._bufs.pop(key, None)
        return self

    def __call__(self, *args):
        if len(args) != len(self.static_args):
            raise RuntimeError(f"GraphedCall was captured with {len(self.static_args)} arguments, got {len(args)}")
        for s, a in zip(self.static_args, args):
            if torch.is_tensor(s):
                if a.shape != s.shape or a.dtype != s.dtype or a.device != s.device:
                    raise RuntimeError(f"GraphedCall: argument {tuple(a.shape)} {a.dtype} does not match the captured "
                                       f"{tuple(s.shape)} {s.dtype}")
                if a.data_ptr() != s.data_ptr():
                    s.copy_(a)
            elif s != a:
                raise RuntimeError(f"GraphedCall: non-tensor argument {a!r} differs from the captured {s!r}")
        self.graph.replay()
        return self.static_out

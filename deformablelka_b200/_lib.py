"""ctypes binding of libdlka_b200.so (the C ABI declared in include/dlka.h).

There is no CPU path and no fallback: if the shared library is missing this module raises at
import, and every compute call on a CPU tensor raises RuntimeError (the reference's 3D op does
the same: "Not implemented on the CPU", 3D/dcn/src/deform_conv.h:46).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_int, c_int32, c_size_t, c_uint64, c_void_p
from typing import Optional

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
# DLKA_LIB selects an experimental build variant of the same library (never a different backend)
LIB_PATH = os.environ.get("DLKA_LIB") or os.path.join(HERE, "libdlka_b200.so")

MATH_FP32_SIMT = 0
MATH_BF16X3 = 1
_MATH_NAMES = {"fp32": MATH_FP32_SIMT, "fp32_simt": MATH_FP32_SIMT, "bf16x3": MATH_BF16X3}


class DwGeom3d(Structure):
    """dlkaDwGeom3d: depthwise stencil shapes of the 3D block (include/dlka.h)."""
    _fields_ = [("conv0_k", c_int * 3), ("conv0_dil", c_int * 3), ("conv_spatial_k", c_int * 3), ("conv_spatial_dil", c_int * 3)]


class Block3dParams(Structure):
    _fields_ = [(n, c_void_p) for n in (
        "proj_1_weight", "proj_1_bias", "conv0_weight", "conv0_bias", "conv_spatial_weight", "conv_spatial_bias",
        "conv_offset_weight", "conv_offset_bias", "deform_weight", "deform_bias", "conv1_weight", "conv1_bias",
        "proj_2_weight", "proj_2_bias")] + [("dw_geom", POINTER(DwGeom3d))]


class Block2dParams(Structure):
    _fields_ = [(n, c_void_p) for n in (
        "proj_1_weight", "proj_1_bias", "conv0_offset_weight", "conv0_offset_bias", "conv0_deform_weight",
        "conv_spatial_offset_weight", "conv_spatial_offset_bias", "conv_spatial_deform_weight",
        "conv1_weight", "conv1_bias", "proj_2_weight", "proj_2_bias")]


class LkaBlock2dParams(Structure):
    _fields_ = ([("norm1_weight", c_void_p), ("norm1_bias", c_void_p), ("attn", Block2dParams), ("layer_scale_1", c_void_p),
                 ("norm2_weight", c_void_p), ("norm2_bias", c_void_p), ("fc1_weight", c_void_p), ("fc1_bias", c_void_p),
                 ("dw_weight", c_void_p), ("dw_bias", c_void_p), ("fc2_weight", c_void_p), ("fc2_bias", c_void_p),
                 ("layer_scale_2", c_void_p), ("eps1", ctypes.c_float), ("eps2", ctypes.c_float), ("hidden", c_int)])


class Transformer3dParams(Structure):
    _fields_ = ([("attn", Block3dParams)] + [(n, c_void_p) for n in (
        "norm_weight", "norm_bias", "gamma", "pos_embed", "conv1_weight", "bn1_scale", "bn1_shift", "conv2_weight",
        "bn2_scale", "bn2_shift", "conv8_weight", "conv8_bias")] + [("eps", ctypes.c_float), ("lrelu_slope", ctypes.c_float)])


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python deformablelka_b200/build.py` "
            "(deformablelka_b200 has no CPU / PyTorch fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    I = c_int
    V = c_void_p
    lib.dlka_version.restype = c_int
    lib.dlka_status_string.restype = c_char_p
    lib.dlka_status_string.argtypes = [c_int]
    lib.dlka_last_cuda_error.restype = c_char_p
    lib.dlka_launch_count.restype = c_uint64
    lib.dlka_profile_enable.restype = c_int
    lib.dlka_profile_enable.argtypes = [c_int]
    lib.dlka_profile_summary.restype = c_int
    lib.dlka_profile_summary.argtypes = [c_char_p, c_size_t]
    lib.dlka_deform_conv3d_workspace_bytes.restype = c_size_t
    lib.dlka_deform_conv3d_workspace_bytes.argtypes = [I] * 20
    lib.dlka_deform_conv3d_forward.restype = c_int
    lib.dlka_deform_conv3d_forward.argtypes = [V] * 5 + [I] * 22 + [V, c_size_t, V]
    lib.dlka_deform_conv3d_sample_indices.restype = c_int
    lib.dlka_deform_conv3d_sample_indices.argtypes = [V] * 3 + [I] * 17 + [V]
    lib.dlka_deform_conv2d_workspace_bytes.restype = c_size_t
    lib.dlka_deform_conv2d_workspace_bytes.argtypes = [I] * 15
    lib.dlka_deform_conv2d_forward.restype = c_int
    lib.dlka_deform_conv2d_forward.argtypes = [V] * 6 + [I] * 16 + [V, c_size_t, V]
    lib.dlka_deform_conv2d_backward_workspace_bytes.restype = c_size_t
    lib.dlka_deform_conv2d_backward_workspace_bytes.argtypes = [I] * 16
    lib.dlka_deform_conv2d_backward.restype = c_int
    lib.dlka_deform_conv2d_backward.argtypes = [V] * 10 + [I] * 15 + [V, c_size_t, V]
    lib.dlka_deform_conv2d_sample_indices.restype = c_int
    lib.dlka_deform_conv2d_sample_indices.argtypes = [V] * 3 + [I] * 12 + [V]
    lib.dlka_deform_conv_pack3d_workspace_bytes.restype = c_size_t
    lib.dlka_deform_conv_pack3d_workspace_bytes.argtypes = [I] * 20
    lib.dlka_deform_conv_pack3d_forward.restype = c_int
    lib.dlka_deform_conv_pack3d_forward.argtypes = [V] * 6 + [I] * 22 + [V, c_size_t, V]
    lib.dlka_deform_conv_pack2d_workspace_bytes.restype = c_size_t
    lib.dlka_deform_conv_pack2d_workspace_bytes.argtypes = [I] * 14
    lib.dlka_deform_conv_pack2d_forward.restype = c_int
    lib.dlka_deform_conv_pack2d_forward.argtypes = [V] * 6 + [I] * 15 + [V, c_size_t, V]
    for name in ("dlka_lka3d_deform", "dlka_lka_attention3d_deform"):
        getattr(lib, name + "_workspace_bytes").restype = c_size_t
        getattr(lib, name + "_workspace_bytes").argtypes = [I] * 5
        getattr(lib, name + "_forward").restype = c_int
        getattr(lib, name + "_forward").argtypes = [POINTER(Block3dParams), V, V] + [I] * 6 + [V, c_size_t, V]
    lib.dlka_lka_attention3d_deform_packed_bytes.restype = c_size_t
    lib.dlka_lka_attention3d_deform_packed_bytes.argtypes = [I]
    lib.dlka_lka_attention3d_deform_forward_packed.restype = c_int
    lib.dlka_lka_attention3d_deform_forward_packed.argtypes = [POINTER(Block3dParams), V, V] + [I] * 6 + [V, c_size_t, I, V, c_size_t, V]
    lib.dlka_deformable_lka_attention2d_packed_bytes.restype = c_size_t
    lib.dlka_deformable_lka_attention2d_packed_bytes.argtypes = [I]
    lib.dlka_deformable_lka_attention2d_forward_packed.restype = c_int
    lib.dlka_deformable_lka_attention2d_forward_packed.argtypes = [POINTER(Block2dParams), V, V] + [I] * 5 + [V, c_size_t, I, V, c_size_t, V]
    lib.dlka_lka_attention3d_deform_forward_host.restype = c_int
    lib.dlka_lka_attention3d_deform_forward_host.argtypes = (
        [POINTER(Block3dParams), V, V] + [I] * 6 + [V, c_size_t, V, c_size_t, V])
    lib.dlka_deformable_lka_block2d_workspace_bytes.restype = c_size_t
    lib.dlka_deformable_lka_block2d_workspace_bytes.argtypes = [I] * 5
    lib.dlka_deformable_lka_block2d_forward.restype = c_int
    lib.dlka_deformable_lka_block2d_forward.argtypes = [POINTER(LkaBlock2dParams), V, V] + [I] * 5 + [V, c_size_t, V]
    lib.dlka_lka_transformer3d_prenorm_workspace_bytes.restype = c_size_t
    lib.dlka_lka_transformer3d_prenorm_workspace_bytes.argtypes = [I] * 5
    lib.dlka_lka_transformer3d_prenorm_forward.restype = c_int
    lib.dlka_lka_transformer3d_prenorm_forward.argtypes = (
        [POINTER(Block3dParams), V, V, ctypes.c_float, V, V, V, V] + [I] * 6 + [V, c_size_t, V])
    lib.dlka_lka_transformer3d_block_workspace_bytes.restype = c_size_t
    lib.dlka_lka_transformer3d_block_workspace_bytes.argtypes = [I] * 5
    lib.dlka_lka_transformer3d_block_forward.restype = c_int
    lib.dlka_lka_transformer3d_block_forward.argtypes = [POINTER(Transformer3dParams), V, V] + [I] * 6 + [V, c_size_t, V]
    lib.dlka_deform_conv3d_backward_workspace_bytes.restype = c_size_t
    lib.dlka_deform_conv3d_backward_workspace_bytes.argtypes = [I] * 20
    lib.dlka_deform_conv3d_backward.restype = c_int
    lib.dlka_deform_conv3d_backward.argtypes = [V] * 8 + [I] * 22 + [V, c_size_t, V]
    lib.dlka_linear_tokens_workspace_bytes.restype = c_size_t
    lib.dlka_linear_tokens_workspace_bytes.argtypes = [I, I]
    lib.dlka_linear_tokens_forward.restype = c_int
    lib.dlka_linear_tokens_forward.argtypes = [V] * 5 + [ctypes.c_longlong, I, I, I, V, c_size_t, V]
    lib.dlka_patch_expand2d_workspace_bytes.restype = c_size_t
    lib.dlka_patch_expand2d_workspace_bytes.argtypes = [I] * 5
    lib.dlka_patch_expand2d_forward.restype = c_int
    lib.dlka_patch_expand2d_forward.argtypes = [V] * 4 + [ctypes.c_float, V] + [I] * 6 + [V, c_size_t, V]
    lib.dlka_host_pipe_create.restype = c_int
    lib.dlka_host_pipe_create.argtypes = [POINTER(c_void_p), c_int]
    for name in ("dlka_host_pipe_destroy", "dlka_host_pipe_wait"):
        getattr(lib, name).restype = c_int
        getattr(lib, name).argtypes = [c_void_p]
    lib.dlka_host_pipe_join.restype = c_int
    lib.dlka_host_pipe_join.argtypes = [c_void_p, c_void_p]
    lib.dlka_host_pipe_completed.restype = ctypes.c_longlong
    lib.dlka_host_pipe_completed.argtypes = [c_void_p]
    lib.dlka_host_numa_node.restype = c_int
    lib.dlka_host_numa_node.argtypes = [c_int]
    lib.dlka_host_bind_thread.restype = c_int
    lib.dlka_host_bind_thread.argtypes = [c_int]
    lib.dlka_host_alloc.restype = c_int
    lib.dlka_host_alloc.argtypes = [POINTER(c_void_p), c_size_t, c_int, c_int]
    lib.dlka_host_free.restype = c_int
    lib.dlka_host_free.argtypes = [c_void_p]
    lib.dlka_lka_attention3d_deform_forward_host_async.restype = c_int
    lib.dlka_lka_attention3d_deform_forward_host_async.argtypes = (
        [c_void_p, POINTER(Block3dParams), V, V] + [I] * 6 + [V, c_size_t, V, c_size_t, V])
    for name in ("dlka_deformable_lka2d", "dlka_deformable_lka_attention2d"):
        getattr(lib, name + "_workspace_bytes").restype = c_size_t
        getattr(lib, name + "_workspace_bytes").argtypes = [I] * 4
        getattr(lib, name + "_forward").restype = c_int
        getattr(lib, name + "_forward").argtypes = [POINTER(Block2dParams), V, V] + [I] * 5 + [V, c_size_t, V]
    return lib


lib = _load()


def math_mode(name_or_int) -> int:
    if isinstance(name_or_int, int):
        return name_or_int
    try:
        return _MATH_NAMES[str(name_or_int).lower()]
    except KeyError:
        raise ValueError(f"unknown math mode {name_or_int!r}; expected one of {sorted(_MATH_NAMES)}")


def default_math() -> int:
    return math_mode(os.environ.get("DLKA_MATH", "bf16x3"))  # product default: tcgen05 path; "fp32" = exact SIMT validation path


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib.dlka_status_string(status).decode()
        cuda = lib.dlka_last_cuda_error().decode()
        raise RuntimeError(f"{what}: {msg}" + (f" [{cuda}]" if cuda and status in (-4, -5) else ""))


def dptr(t: Optional[torch.Tensor], what: str = "tensor") -> Optional[int]:
    """Device pointer of a contiguous fp32 CUDA tensor (None passes through as NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"{what}: Not implemented on the CPU (deformablelka_b200 is CUDA-only)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{what}: expected float32, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{what} tensor has to be contiguous")
    return t.data_ptr()


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class Workspace:
    """Per-device grow-only scratch buffer handed to the library (it never allocates itself)."""

    _bufs = {}

    @classmethod
    def get(cls, device: torch.device, nbytes: int) -> torch.Tensor:
        key = (device.index if device.index is not None else torch.cuda.current_device(),
               torch.cuda.current_stream(device).cuda_stream)
        buf = cls._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            cls._bufs[key] = None
            buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
            cls._bufs[key] = buf
        return buf

    @classmethod
    def clear(cls) -> None:
        cls._bufs.clear()


def launch_count() -> int:
    return int(lib.dlka_launch_count())


def profile_enable(on: bool) -> None:
    lib.dlka_profile_enable(1 if on else 0)


def profile_summary() -> dict:
    """{kernel name: (launches, total_ms)} since the last call; synchronises the recorded events."""
    buf = ctypes.create_string_buffer(1 << 16)
    check(lib.dlka_profile_summary(buf, len(buf)), "dlka_profile_summary")
    out = {}
    for line in buf.value.decode().splitlines():
        name, n, ms = line.split()
        out[name] = (int(n), float(ms))
    return out

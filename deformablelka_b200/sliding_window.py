"""Sliding-window 3D inference (rest of SURVEY.md 8f row N4): tiled prediction of a whole volume with Gaussian-weighted
overlap and mirroring test-time augmentation, the way the reference's ``SegmentationNetwork`` feeds real volumes through
the D-LKA network (3D/d_lka_former/network_architecture/neural_network.py:250-428, 502-556).

Host-side orchestration only: the network is any callable ``[1, c, px, py, pz] -> [1, num_classes, px, py, pz]`` (logits);
aggregation stays on the network's device in fp32 (the reference's ``all_in_gpu`` path aggregates in fp16, its default path on
the host in fp32).  ``batchgenerators.pad_nd_image`` (not vendored in the reference) is restated from its published
behaviour: symmetric constant padding up to the patch size, split ``d // 2`` below and the rest above."""
from __future__ import annotations

import itertools
from typing import Callable, List, Sequence, Tuple

import numpy as np
import torch


def compute_steps_for_sliding_window(patch_size: Sequence[int], image_size: Sequence[int], step_size: float) -> List[List[int]]:
    """Window origins per axis (neural_network.py:267-290): at most ``patch * step_size`` apart, evenly spread so the last
    window ends exactly at the image border."""
    assert all(i >= p for i, p in zip(image_size, patch_size)), "image size must be as large or larger than patch_size"
    assert 0 < step_size <= 1, "step_size must be larger than 0 and smaller or equal to 1"
    steps = []
    for img, patch in zip(image_size, patch_size):
        n = int(np.ceil((img - patch) / (patch * step_size))) + 1
        span = img - patch
        actual = span / (n - 1) if n > 1 else 0.0
        steps.append([int(np.round(actual * i)) for i in range(n)])
    return steps


def gaussian_importance_map(patch_size: Sequence[int], sigma_scale: float = 1.0 / 8) -> np.ndarray:
    """Centre-peaked weight of a patch (neural_network.py:251-264): unit impulse at the centre, Gaussian-filtered with
    sigma = patch * sigma_scale, scaled to max 1, zeros lifted to the smallest positive value (no 0/0 later)."""
    from scipy.ndimage import gaussian_filter
    tmp = np.zeros(tuple(patch_size))
    tmp[tuple(p // 2 for p in patch_size)] = 1
    g = gaussian_filter(tmp, [p * sigma_scale for p in patch_size], 0, mode="constant", cval=0)
    g = (g / g.max()).astype(np.float32)
    g[g == 0] = g[g != 0].min()
    return g


def pad_to_patch(x: torch.Tensor, patch_size: Sequence[int]) -> Tuple[torch.Tensor, Tuple[slice, ...]]:
    """Zero-pad the three trailing axes of ``x`` up to ``patch_size`` (symmetric, remainder above); returns the padded tensor
    and the slices that undo the padding."""
    pads, crop = [], []
    for n, p in zip(x.shape[-3:], patch_size):
        d = max(p - n, 0)
        pads.append((d // 2, d - d // 2))
        crop.append(slice(d // 2, d // 2 + n))
    flat = [v for lo_hi in reversed(pads) for v in lo_hi]     # F.pad wants the last axis first
    return torch.nn.functional.pad(x, flat), tuple(crop)


def mirror_and_predict(network: Callable, x: torch.Tensor, mirror_axes: Sequence[int] = (0, 1, 2), do_mirroring: bool = True,
                       mult: torch.Tensor = None, nonlin: Callable = lambda t: torch.softmax(t, 1)) -> torch.Tensor:
    """Average of ``nonlin(network(flip(x)))`` flipped back over every subset of ``mirror_axes`` (neural_network.py:502-556);
    the sum is divided by ``2 ** len(mirror_axes)`` exactly as the reference does."""
    assert x.dim() == 5, "x must be (b, c, x, y, z)"
    subsets = [()]
    if do_mirroring:
        subsets = [s for r in range(len(mirror_axes) + 1) for s in itertools.combinations(sorted(mirror_axes), r)]
    scale = 1.0 / (2 ** len(mirror_axes) if do_mirroring else 1)
    out = None
    for s in subsets:
        dims = tuple(a + 2 for a in s)
        pred = nonlin(network(torch.flip(x, dims) if dims else x))
        pred = torch.flip(pred, dims) if dims else pred
        out = pred * scale if out is None else out + pred * scale
    if mult is not None:
        out = out * mult
    return out


def predict_3d_tiled(network: Callable, x, patch_size: Sequence[int], num_classes: int, step_size: float = 0.5,
                     do_mirroring: bool = True, mirror_axes: Sequence[int] = (0, 1, 2), use_gaussian: bool = True,
                     regions_class_order=None, device=None):
    """Tiled prediction of a volume ``x`` (c, X, Y, Z) -> (segmentation (X, Y, Z), class probabilities (classes, X, Y, Z)),
    numpy arrays, following ``_internal_predict_3D_3Dconv_tiled`` (neural_network.py:292-428)."""
    x = torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x, dtype=torch.float32)
    assert x.dim() == 4, "x must be (c, x, y, z)"
    assert patch_size is not None, "patch_size cannot be None for tiled prediction"
    if device is not None:
        x = x.to(device)
    data, crop = pad_to_patch(x, patch_size)
    steps = compute_steps_for_sliding_window(patch_size, data.shape[1:], step_size)
    num_tiles = len(steps[0]) * len(steps[1]) * len(steps[2])
    gauss = None
    if use_gaussian and num_tiles > 1:
        gauss = torch.from_numpy(gaussian_importance_map(patch_size)).to(data.device)
    add = gauss if gauss is not None else torch.ones(tuple(patch_size), device=data.device)
    agg = torch.zeros((num_classes,) + tuple(data.shape[1:]), device=data.device)
    cnt = torch.zeros_like(agg)
    px, py, pz = patch_size
    with torch.no_grad():
        for lx in steps[0]:
            for ly in steps[1]:
                for lz in steps[2]:
                    patch = data[None, :, lx:lx + px, ly:ly + py, lz:lz + pz]
                    pred = mirror_and_predict(network, patch, mirror_axes, do_mirroring, gauss)[0]
                    agg[:, lx:lx + px, ly:ly + py, lz:lz + pz] += pred
                    cnt[:, lx:lx + px, ly:ly + py, lz:lz + pz] += add
    sl = (slice(None),) + crop
    probs = (agg[sl] / cnt[sl]).cpu().numpy()
    if regions_class_order is None:
        seg = probs.argmax(0)
    else:
        seg = np.zeros(probs.shape[1:], dtype=np.float32)
        for i, c in enumerate(regions_class_order):
            seg[probs[i] > 0.5] = c
    return seg, probs

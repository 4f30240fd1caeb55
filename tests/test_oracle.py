"""Pins for the CPU oracle (runs without a GPU).

K1 zero offsets == stock conv; K2 3D op with D=1 == torchvision deform_conv2d;
K3 fresh DeformConvPack == nn.Conv3d; golden vectors from the unmodified reference 2D module.
"""
import glob
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F
import torchvision

from conftest import GOLDEN


def _load_golden(path):
    z = np.load(path)
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    return torch.from_numpy(z["x"]), torch.from_numpy(z["y"]), sd


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "ref2d_*_s[0-9].npz"))), ids=os.path.basename)
@pytest.mark.parametrize("impl", ["torchvision", "c"])
def test_golden_2d_reference_module(oracle, path, impl):
    x, y, sd = _load_golden(path)
    dim = x.shape[1]
    cls = oracle.deformable_LKA if "_lka_" in os.path.basename(path) else oracle.deformable_LKA_Attention
    m = cls(dim, impl=impl).eval()
    missing = m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        got = m(x)
    tol = 0 if impl == "torchvision" else 2e-5
    assert (got - y).abs().max().item() <= tol * max(1.0, y.abs().max().item())


@pytest.mark.parametrize("groups,wg,mask", [(1, 1, False), (2, 4, True), (1, 8, False)])
def test_c_deform_conv2d_vs_torchvision(oracle, groups, wg, mask):
    torch.manual_seed(0)
    B, C, H, W, Co, kh, kw = 2, 8, 11, 9, 8, 3, 5
    x = torch.randn(B, C, H, W)
    stride, pad, dil = (1, 2), (2, 3), (2, 1)
    Ho = oracle.out_extent(H, pad[0], dil[0], kh, stride[0]); Wo = oracle.out_extent(W, pad[1], dil[1], kw, stride[1])
    off = torch.randn(B, groups * 2 * kh * kw, Ho, Wo) * 3
    m = torch.rand(B, groups * kh * kw, Ho, Wo) if mask else None
    w = torch.randn(Co, C // wg, kh, kw); b = torch.randn(Co)
    ref = torchvision.ops.deform_conv2d(x, off, w, b, stride, pad, dil, m)
    got = oracle.deform_conv2d_c(x, off, w, b, stride, pad, dil, m)
    assert torch.allclose(got, ref, atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("groups,stride,pad,dil,k", [(1, 1, 1, 1, 3), (4, 1, 2, 1, 5), (8, 1, 9, 3, 7), (2, (1, 2, 1), (0, 1, 2), (1, 2, 1), (1, 3, 2))])
def test_k1_zero_offset_equals_conv3d(oracle, groups, stride, pad, dil, k):
    torch.manual_seed(1)
    B, C, Co = 2, 8, 8
    D, H, W = (7, 9, 8) if k != 7 else (20, 21, 22)
    kd, kh, kw = oracle._triple(k)
    x = torch.randn(B, C, D, H, W)
    w = torch.randn(Co, C // groups, kd, kh, kw); b = torch.randn(Co)
    ref = F.conv3d(x, w, b, stride, pad, dil, groups)
    off = torch.zeros(B, 3 * kd * kh * kw, *ref.shape[2:])
    got = oracle.deform_conv3d(x, off, w, b, stride, pad, dil, groups, 1)
    assert torch.allclose(got, ref, atol=2e-4, rtol=1e-4)
    got_c = oracle.deform_conv3d_c(x, off, w, b, stride, pad, dil, groups, 1)
    assert torch.allclose(got_c, ref, atol=2e-4, rtol=1e-4)


def test_k2_3d_with_depth1_equals_torchvision_2d(oracle):
    torch.manual_seed(2)
    B, C, H, W, Co, kh, kw = 2, 6, 10, 12, 4, 3, 3
    x = torch.randn(B, C, H, W)
    off2 = torch.randn(B, 2 * kh * kw, H, W) * 4  # many samples out of bounds
    w = torch.randn(Co, C, kh, kw); b = torch.randn(Co)
    ref = torchvision.ops.deform_conv2d(x, off2, w, b, 1, 1, 1)
    off3 = torch.zeros(B, 3 * kh * kw, 1, H, W)
    off3[:, 1::3] = off2[:, 0::2].unsqueeze(2)
    off3[:, 2::3] = off2[:, 1::2].unsqueeze(2)
    got = oracle.deform_conv3d(x.unsqueeze(2), off3, w.unsqueeze(2), b, 1, (0, 1, 1), 1, 1, 1)
    assert torch.allclose(got.squeeze(2), ref, atol=1e-4, rtol=1e-4)


def test_k2_deformable_groups(oracle):
    torch.manual_seed(3)
    B, C, H, W, Co, kh, kw, dg = 1, 8, 9, 7, 8, 3, 3, 2
    x = torch.randn(B, C, H, W)
    off2 = torch.randn(B, dg * 2 * kh * kw, H, W) * 2
    w = torch.randn(Co, C // 2, kh, kw); b = torch.randn(Co)
    ref = torchvision.ops.deform_conv2d(x, off2, w, b, 1, 1, 1)
    off3 = torch.zeros(B, dg * 3 * kh * kw, 1, H, W)
    o2 = off2.view(B, dg, kh * kw, 2, H, W)
    o3 = off3.view(B, dg, kh * kw, 3, 1, H, W)
    o3[:, :, :, 1, 0] = o2[:, :, :, 0]; o3[:, :, :, 2, 0] = o2[:, :, :, 1]
    got = oracle.deform_conv3d(x.unsqueeze(2), off3, w.unsqueeze(2), b, 1, (0, 1, 1), 1, 2, dg)
    assert torch.allclose(got.squeeze(2), ref, atol=1e-4, rtol=1e-4)


def test_k3_fresh_pack_equals_conv3d(oracle):
    torch.manual_seed(4)
    m = oracle.DeformConvPack3D(6, 6, (3, 3, 3), 1, 1).eval()
    x = torch.randn(2, 6, 5, 6, 7)
    with torch.no_grad():
        ref = F.conv3d(x, m.weight, m.bias, 1, 1)
        got = m(x)
    assert torch.allclose(got, ref, atol=1e-4, rtol=1e-4)


def test_addmm_and_c_gemm_variants_agree(oracle):
    torch.manual_seed(5)
    x = torch.randn(2, 8, 6, 7, 5)
    off = torch.randn(2, 81, 6, 7, 5) * 2
    w = torch.randn(8, 8, 3, 3, 3); b = torch.randn(8)
    a = oracle.deform_conv3d(x, off, w, b, 1, 1, 1, chunk=50)
    c = oracle.deform_conv3d_c(x, off, w, b, 1, 1, 1)
    assert torch.allclose(a, c, atol=1e-4, rtol=1e-4)


def test_sample_indices_3d_consistent(oracle):
    torch.manual_seed(6)
    D, H, W = 4, 5, 6
    off = torch.randn(1, 81, D, H, W) * 3
    low, mask = oracle.sample_indices3d(off, (D, H, W), 3, 1, 1, 1)
    assert low.shape == (1, D * H * W, 27, 3) and mask.shape == (1, D * H * W, 27)
    # recompute floor in torch
    o = off.view(27, 3, D, H, W)
    dd, hh, ww = torch.meshgrid(torch.arange(D), torch.arange(H), torch.arange(W), indexing="ij")
    t = torch.arange(27)
    base = torch.stack([t // 9 - 1, (t // 3) % 3 - 1, t % 3 - 1], 1).view(27, 3, 1, 1, 1).float()
    p = torch.stack([dd, hh, ww]).float().unsqueeze(0) + base + o
    fl = torch.floor(p).to(torch.int32).permute(2, 3, 4, 0, 1).reshape(D * H * W, 27, 3)
    valid = ((p > -1).all(1) & (p[:, 0] < D) & (p[:, 1] < H) & (p[:, 2] < W)).permute(1, 2, 3, 0).reshape(D * H * W, 27)
    assert torch.equal((mask[0] & 1).bool(), valid)
    assert torch.equal(low[0][valid], fl[valid])


def test_block3d_oracle_runs_and_identity(oracle):
    torch.manual_seed(7)
    m = oracle.LKA_Attention3d_deform(8).eval()
    B, C, H, W, D = 1, 8, 6, 5, 4
    x = torch.randn(B, H * W * D, C)
    with torch.no_grad():
        y0 = m(x, B, C, H, W, D)
        # zero-init conv_offset => block equals the same block with a plain Conv3d (K3 at block level)
        sg = m.spatial_gating_unit
        xx = x.permute(0, 2, 1).reshape(B, C, H, W, D)
        t = F.gelu(m.proj_1(xx))
        a = F.conv3d(sg.conv_spatial(sg.conv0(t)), sg.deform_conv.weight, sg.deform_conv.bias, 1, 1)
        ref = m.proj_2(t * sg.conv1(a)) + xx
    assert torch.allclose(y0, ref.reshape(B, C, -1).permute(0, 2, 1), atol=1e-4, rtol=1e-4)
    oracle.randomize_offsets_(m)
    with torch.no_grad():
        y1 = m(x, B, C, H, W, D)
    assert (y1 - y0).abs().max() > 1e-3

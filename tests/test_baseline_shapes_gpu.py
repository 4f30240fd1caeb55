"""Parity on the configurations that are benchmarked (BASELINE.json configs C1-C4 and the headline shape H), with
non-trivial offsets everywhere:

* C1   1x64x224x224 2D block: product vs the golden vectors of the UNMODIFIED reference module (stored on a lattice, the
       input regenerated from its seed) and vs the oracle on the full tensor;
* C2   the three 2D-net block shapes [24,384,14,14], [24,192,28,28], [24,96,56,56] vs the oracle;
* C4   the 3D-net block shapes (2,32,32^3), (2,64,16^3), (2,128,8^3), (2,256,4^3) vs the oracle;
* mid  2x96x24x40x48 with random conv_offset weights vs the full oracle (multi-tile zero-copy conv, lattice stencils and
       deformable bricks all compared with real offsets);
* H    (2,96,64,128,128) with random offsets: the oracle is run on crop + halo regions (halo = 11 stencil + 1 offset conv +
       1 tap + ceil(max|offset|) + 1 corner) and compared on the crop interior, for crops at both volume corners, in the
       interior, on a face, and at the highest addresses of sample 1;
* the sliding-window predictor on the GPU with a native D-LKA block inside, vs the numpy restatement of the reference's tiled
  prediction run with the oracle block.

Tolerance 1e-3 of the output range (north_star); (C3 is tests/test_ref_d3d_gpu.py: vs the reference's own compiled D3D).
"""
import math as pymath
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
TOL = 1e-3
DEV = "cuda:0"


def rel_err(got, ref):
    got = got.detach().float().cpu(); ref = ref.detach().float().cpu()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def dl():
    import deformablelka_b200 as d
    return d


@pytest.fixture(params=["bf16x3", "fp32"])
def math(request, monkeypatch):
    monkeypatch.setenv("DLKA_MATH", request.param)
    return request.param


@pytest.fixture
def tc_math(monkeypatch):
    monkeypatch.setenv("DLKA_MATH", "bf16x3")   # the benchmarked arithmetic


# ------------------------------------------------------------------------------------------ C1
def test_c1_block2d_vs_reference_golden_and_oracle(dl, oracle, math):
    z = np.load(os.path.join(GOLDEN, "ref2d_c1_attn_c64.npz"))
    sd = {k[3:]: torch.from_numpy(z[k]).float() for k in z.files if k.startswith("sd.")}
    x = torch.randn(1, 64, 224, 224, generator=torch.Generator().manual_seed(int(z["meta.x_seed"])))
    assert torch.equal(x.flatten()[:16], torch.from_numpy(z["meta.x_head"]))                 # same input as the reference saw
    assert abs(float(x.double().sum()) - float(z["meta.x_sum"])) < 1e-6 * float(z["meta.x_abs_sum"])
    m = dl.deformable_LKA_Attention(64)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval()
    with torch.no_grad():
        y = m(x.to(DEV)).cpu()
        y_lka = m.spatial_gating_unit(x.to(DEV)).cpu()
    for got, key in ((y, "y"), (y_lka, "y_lka")):
        ref_sub = torch.from_numpy(z[f"out.{key}_sub"])
        scale = ref_sub.abs().max()
        assert ((got[:, :, ::5, ::3] - ref_sub).abs().max() / scale).item() < TOL
        chan = got.double().sum((0, 2, 3))
        ref_chan = torch.from_numpy(z[f"out.{key}_chan_sum"])
        assert ((chan - ref_chan).abs().max() / (224 * 224 * scale)).item() < 1e-5          # mean error per pixel, full tensor
    om = oracle.deformable_LKA_Attention(64).eval()
    om.load_state_dict(sd, strict=True)
    with torch.no_grad():
        assert rel_err(y, om(x)) < TOL


# ------------------------------------------------------------------------------------------ C2
@pytest.mark.parametrize("B,C,H,W", [(24, 384, 14, 14), (24, 192, 28, 28), (24, 96, 56, 56)])
def test_c2_block2d_shapes_vs_oracle(dl, oracle, B, C, H, W, tc_math):
    torch.manual_seed(50 + C)
    ref_m = oracle.deformable_LKA_Attention(C).eval()       # default init: mean |offset| ~ 0.2 .. 0.5 px
    m = dl.deformable_LKA_Attention(C)
    m.load_state_dict(ref_m.state_dict())
    x = torch.randn(B, C, H, W)
    with torch.no_grad():
        ref = ref_m(x)
        got = m.to(DEV)(x.to(DEV))
    assert rel_err(got, ref) < TOL


# ------------------------------------------------------------------------------------------ C4
@pytest.mark.parametrize("C,S", [(32, 32), (64, 16), (128, 8), (256, 4)])
def test_c4_block3d_shapes_vs_oracle(dl, oracle, C, S, math):
    torch.manual_seed(60 + C)
    B = 2
    ref_m = oracle.LKA_Attention3d_deform(C).eval()
    oracle.randomize_offsets_(ref_m, std=0.05, bias_range=1.0)
    m = dl.LKA_Attention3d_deform(C)
    m.load_state_dict(ref_m.state_dict())
    x = torch.randn(B, S * S * S, C)
    with torch.no_grad():
        ref = ref_m(x, B, C, S, S, S)
        got = m.to(DEV)(x.to(DEV), B, C, S, S, S)
    assert rel_err(got, ref) < TOL


# ------------------------------------------------------------------------------------------ mid-size, real offsets
def test_block3d_midsize_random_offsets_vs_full_oracle(dl, oracle, math):
    torch.manual_seed(70)
    B, C, dims = 2, 96, (24, 40, 48)
    ref_m = oracle.LKA_Attention3d_deform(C).eval()
    oracle.randomize_offsets_(ref_m, std=0.05, bias_range=1.0)
    m = dl.LKA_Attention3d_deform(C)
    m.load_state_dict(ref_m.state_dict())
    m = m.to(DEV)
    N = dims[0] * dims[1] * dims[2]
    x = torch.randn(B, N, C)
    with torch.no_grad():
        ref = ref_m(x, B, C, *dims)
        got = m(x.to(DEV), B, C, *dims)
        xv = x.permute(0, 2, 1).reshape(B, C, *dims).contiguous()
        ref_l = ref_m.spatial_gating_unit(xv)
        got_l = m.spatial_gating_unit(xv.to(DEV))
    assert rel_err(got, ref) < TOL
    assert rel_err(got_l, ref_l) < TOL


# ------------------------------------------------------------------------------------------ headline shape, crops
def _oracle_offsets_absmax(ref_m, xc, interior):
    """max |offset| over the crop interior (exact there once the halo covers the stencils + the offset conv)."""
    sg = ref_m.spatial_gating_unit
    with torch.no_grad():
        a = sg.conv_spatial(sg.conv0(ref_m.activation(ref_m.proj_1(xc))))
        off = sg.deform_conv.conv_offset(a)
    return off[(slice(None), slice(None)) + interior].abs().max().item()


def test_headline_random_offsets_vs_oracle_crops(dl, oracle, tc_math):
    B, C, dims = 2, 96, (64, 128, 128)
    I, HALO = 8, 19                      # 8^3 interior; halo 19 = 11 + 1 + 1 + 5 + 1  ->  |offset| up to 5 voxels
    torch.manual_seed(80)
    ref_m = oracle.LKA_Attention3d_deform(C).eval()
    oracle.randomize_offsets_(ref_m, std=0.05, bias_range=1.0)
    m = dl.LKA_Attention3d_deform(C)
    m.load_state_dict(ref_m.state_dict())
    m = m.to(DEV)
    N = dims[0] * dims[1] * dims[2]
    x = torch.randn(B, N, C, device=DEV)
    with torch.no_grad():
        y = m(x, B, C, *dims)
    xv = x.view(B, *dims, C)
    yv = y.view(B, *dims, C)
    crops = [
        (0, (0, 0, 0)),                                   # volume corner: zero padding on three faces
        (0, (28, 60, 60)),                                # interior
        (0, (30, 0, 120)),                                # one low face + one high face
        (1, (24, 90, 33)),                                # interior of the second sample
        (1, tuple(d - I for d in dims)),                  # far corner of sample 1: the highest addresses of every tensor
    ]
    worst = 0.0
    for b, start in crops:
        lo = [max(s - HALO, 0) for s in start]
        hi = [min(s + I + HALO, d) for s, d in zip(start, dims)]
        cd = [h - l for l, h in zip(lo, hi)]
        xc_tok = xv[b, lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2], :].reshape(1, -1, C).cpu()
        interior = tuple(slice(s - l, s - l + I) for s, l in zip(start, lo))
        xc = xc_tok.permute(0, 2, 1).reshape(1, C, *cd)
        amax = _oracle_offsets_absmax(ref_m, xc, interior)
        assert 11 + 1 + 1 + pymath.ceil(amax) + 1 <= HALO, f"offsets up to {amax:.2f} voxels need a larger halo"
        with torch.no_grad():
            ref = ref_m(xc_tok, 1, C, *cd).view(*cd, C)[interior]
        got = yv[b, start[0]:start[0] + I, start[1]:start[1] + I, start[2]:start[2] + I, :].cpu()
        err = rel_err(got, ref)
        worst = max(worst, err)
        assert err < TOL, (b, start, err)
    print(f"headline crops: worst rel err {worst:.2e}")


# ------------------------------------------------------------------------------------------ sliding window on the GPU
class _NativeNet(torch.nn.Module):
    """stem conv -> native TransformerBlock_3D_single_deform_LKA (one library call) -> 1x1x1 head (logits)."""

    def __init__(self, dl, cin, C, classes, patch):
        super().__init__()
        n = patch[0] * patch[1] * patch[2]
        self.stem = torch.nn.Conv3d(cin, C, 3, padding=1)
        self.block = dl.TransformerBlock_3D_single_deform_LKA(n, C, C, 4, pos_embed=True)
        self.head = torch.nn.Conv3d(C, classes, 1)

    def forward(self, x):
        with torch.backends.cudnn.flags(allow_tf32=False):
            return self.head(self.block(self.stem(x)))


def test_predict_3d_tiled_native_block_vs_oracle(dl, oracle, math):
    from deformablelka_b200 import sliding_window as sw
    torch.manual_seed(90)
    cin, C, classes, patch = 2, 32, 4, (8, 8, 8)
    net = _NativeNet(dl, cin, C, classes, patch).eval()
    oracle.randomize_offsets_(net)
    blk = net.block
    with torch.no_grad():
        blk.gamma.uniform_(0.2, 1.0); blk.pos_embed.normal_(0, 0.5)
        for bn in (blk.conv51.norm1, blk.conv51.norm2):
            bn.running_mean.normal_(0, 0.3); bn.running_var.uniform_(0.5, 2.0); bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.2)
    # the same network on the CPU with the oracle's block restatement
    ref_attn = oracle.LKA_Attention3d_deform(C).eval()
    ref_attn.load_state_dict(blk.epa_block.state_dict())
    ref_res = oracle.UnetResBlock3D(C).eval()
    ref_res.load_state_dict(blk.conv51.state_dict())
    stem, head = net.stem, net.head
    norm, gamma, pos, conv8 = blk.norm, blk.gamma.detach().clone(), blk.pos_embed.detach().clone(), blk.conv8[1]

    import copy
    stem_c, head_c, norm_c, conv8_c = (copy.deepcopy(t) for t in (stem, head, norm, conv8))

    def oracle_net(xp):
        with torch.no_grad():
            t = stem_c(xp)
            t = oracle.transformer3d_block(norm_c, gamma, ref_attn, pos, ref_res, conv8_c, t)
            return head_c(t)

    x = torch.randn(cin, 12, 14, 10)
    seg_o, probs_o = oracle.sliding_window_predict_oracle(oracle_net, x.numpy(), patch, classes, 0.5, True, (0, 1, 2), True)
    net = net.to(DEV)
    with torch.no_grad():
        seg, probs = sw.predict_3d_tiled(net, x.to(DEV), patch, classes, 0.5, True, (0, 1, 2), True)
    assert probs.shape == probs_o.shape and seg.shape == seg_o.shape
    assert np.abs(probs - probs_o).max() < TOL            # probabilities live in [0, 1]
    diff = seg != seg_o
    if diff.any():                                        # argmax may flip only where the top two classes are tied
        top2 = np.sort(probs_o, 0)[-2:]
        assert np.all((top2[1] - top2[0])[diff] < 2 * TOL)

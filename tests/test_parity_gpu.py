"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle, the golden vectors of the
unmodified reference 2D module, and size-independent properties at full size.

Tolerance (north_star): outputs within 1e-3 relative fp32, measured as max|got-ref| / max|ref|;
integer sampling planes bit-exact.
"""
import glob
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
TOL = 1e-3
DEV = "cuda:0"


def rel_err(got, ref):
    got = got.detach().float().cpu(); ref = ref.detach().float().cpu()
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def dl():
    import deformablelka_b200 as d
    return d


MATHS = ["fp32", "bf16x3"]


@pytest.fixture(params=MATHS)
def math(request, monkeypatch):
    monkeypatch.setenv("DLKA_MATH", request.param)
    return request.param


# ----------------------------------------------------------------------------- golden (reference 2D module)
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "ref2d_*_s[0-9].npz"))), ids=os.path.basename)
def test_golden_2d(dl, path, math):
    z = np.load(path)
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    x, y = torch.from_numpy(z["x"]), torch.from_numpy(z["y"])
    cls = dl.deformable_LKA if "_lka_" in os.path.basename(path) else dl.deformable_LKA_Attention
    m = cls(x.shape[1])
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval()
    with torch.no_grad():
        got = m(x.to(DEV))
    assert got.shape == y.shape
    assert rel_err(got, y) < TOL


# ----------------------------------------------------------------------------- 2D operator vs torchvision (CPU)
@pytest.mark.parametrize("C,Co,wg,og,k,stride,pad,dil,use_mask,use_bias", [
    (8, 8, 1, 1, (3, 3), 1, 1, 1, False, True),
    (16, 16, 16, 1, (5, 5), 1, 2, 1, False, False),       # depthwise, as conv0
    (16, 16, 16, 1, (7, 7), 1, 9, 3, False, False),       # depthwise dilated, as conv_spatial
    (16, 8, 2, 2, (3, 5), (1, 2), (2, 3), (2, 1), True, True),
    (8, 8, 8, 2, (3, 3), 2, 1, 1, True, True),             # depthwise + 2 offset groups + mask
])
def test_deform_conv2d_vs_torchvision(dl, C, Co, wg, og, k, stride, pad, dil, use_mask, use_bias):
    import torchvision
    torch.manual_seed(0)
    B, H, W = 2, 13, 11
    kh, kw = k
    x = torch.randn(B, C, H, W)
    w = torch.randn(Co, C // wg, kh, kw)
    b = torch.randn(Co) if use_bias else None
    from oracle import oracle as o
    sh, sw = o._pair(stride); ph, pw = o._pair(pad); dh, dw = o._pair(dil)
    Ho, Wo = o.out_extent(H, ph, dh, kh, sh), o.out_extent(W, pw, dw, kw, sw)
    off = torch.randn(B, og * 2 * kh * kw, Ho, Wo) * 3
    mask = torch.rand(B, og * kh * kw, Ho, Wo) if use_mask else None
    ref = torchvision.ops.deform_conv2d(x, off, w, b, stride, pad, dil, mask)
    got = dl.ops.deform_conv2d(x.to(DEV), off.to(DEV), w.to(DEV), None if b is None else b.to(DEV), stride, pad, dil,
                               None if mask is None else mask.to(DEV))
    assert rel_err(got, ref) < TOL


def test_deform_conv2d_module_mirrors_torchvision_module(dl):
    import torchvision
    torch.manual_seed(1)
    ref = torchvision.ops.DeformConv2d(8, 8, 3, padding=1, groups=8, bias=False)
    mine = dl.DeformConv2d(8, 8, 3, padding=1, groups=8, bias=False)
    mine.load_state_dict(ref.state_dict())
    x = torch.randn(1, 8, 9, 9); off = torch.randn(1, 18, 9, 9)
    with torch.no_grad():
        assert rel_err(mine.to(DEV)(x.to(DEV), off.to(DEV)), ref(x, off)) < TOL


# ----------------------------------------------------------------------------- 3D operator vs oracle
@pytest.mark.parametrize("C,Co,g,dg,k,stride,pad,dil,scale", [
    (8, 8, 1, 1, 3, 1, 1, 1, 0.0),      # K1: zero offsets == plain conv
    (8, 8, 1, 1, 3, 1, 1, 1, 1.0),
    (16, 12, 1, 1, 3, 1, 1, 1, 8.0),    # many samples out of bounds
    (16, 16, 2, 2, (3, 2, 3), (1, 2, 1), (1, 0, 2), (1, 2, 1), 2.0),
    (8, 8, 8, 1, 5, 1, 2, 1, 1.5),      # depthwise deformable 3D (3D/dcn/test.py:28)
])
def test_deform_conv3d_vs_oracle(dl, oracle, C, Co, g, dg, k, stride, pad, dil, scale):
    torch.manual_seed(2)
    B, D, H, W = 2, 6, 7, 9
    kd, kh, kw = oracle._triple(k)
    sd, sh, sw = oracle._triple(stride); pd, ph, pw = oracle._triple(pad); dd, dh, dw = oracle._triple(dil)
    Do, Ho, Wo = oracle.out_extent(D, pd, dd, kd, sd), oracle.out_extent(H, ph, dh, kh, sh), oracle.out_extent(W, pw, dw, kw, sw)
    x = torch.randn(B, C, D, H, W)
    w = torch.randn(Co, C // g, kd, kh, kw) * 0.2
    b = torch.randn(Co)
    off = torch.randn(B, dg * 3 * kd * kh * kw, Do, Ho, Wo) * scale
    ref = oracle.deform_conv3d(x, off, w, b, stride, pad, dil, g, dg)
    got = dl.ops.deform_conv3d_forward(x.to(DEV), w.to(DEV), b.to(DEV), off.to(DEV), (kd, kh, kw), stride, pad, dil, g, dg, 64)
    assert got.shape == ref.shape
    assert rel_err(got, ref) < TOL
    if scale == 0.0:
        assert rel_err(got, F.conv3d(x, w, b, stride, pad, dil, g)) < TOL


def test_deform_conv3d_module_and_function(dl, oracle):
    torch.manual_seed(3)
    m = dl.DeformConv3d(8, 8, 3, 1, 1).to(DEV)
    x = torch.randn(1, 8, 5, 6, 7); off = torch.randn(1, 81, 5, 6, 7)
    got = m(x.to(DEV), off.to(DEV))
    ref = oracle.deform_conv3d(x, off, m.weight.detach().cpu(), m.bias.detach().cpu(), 1, 1, 1)
    assert rel_err(got, ref) < TOL
    got2 = dl.DeformConvFunction.apply(x.to(DEV), off.to(DEV), m.weight, m.bias, 1, 1, 1, 1, 1, 64)
    assert torch.equal(got, got2)
    with pytest.raises(RuntimeError, match="contiguous"):
        dl.ops.deform_conv3d_forward(x.to(DEV).transpose(3, 4), m.weight, m.bias, off.to(DEV), 3, 1, 1, 1, 1, 1)
    with pytest.raises(RuntimeError, match="must divide im2col_step"):
        dl.ops.deform_conv3d_forward(x.to(DEV).repeat(3, 1, 1, 1, 1), m.weight, m.bias, off.to(DEV).repeat(3, 1, 1, 1, 1),
                                     3, 1, 1, 1, 1, 1, im2col_step=2)


def test_k3_fresh_pack_equals_conv3d(dl):
    torch.manual_seed(4)
    m = dl.DeformConvPack(8, 8, (3, 3, 3), 1, 1).to(DEV)
    x = torch.randn(2, 8, 5, 6, 7)
    got = m(x.to(DEV))
    ref = F.conv3d(x, m.weight.detach().cpu(), m.bias.detach().cpu(), 1, 1)
    assert rel_err(got, ref) < TOL


def test_pack3d_vs_oracle(dl, oracle):
    torch.manual_seed(5)
    m = dl.DeformConvPack(8, 12, (3, 3, 3), 1, 1)
    oracle.randomize_offsets_(m, std=0.1, bias_range=1.5)
    om = oracle.DeformConvPack3D(8, 12, (3, 3, 3), 1, 1)
    om.load_state_dict(m.state_dict())
    x = torch.randn(2, 8, 6, 5, 7)
    with torch.no_grad():
        ref = om(x)
        got = m.to(DEV)(x.to(DEV))
    assert rel_err(got, ref) < TOL


def test_pack2d_vs_oracle(dl, oracle):
    torch.manual_seed(6)
    m = dl.DeformConv(8, groups=8, kernel_size=(7, 7), padding=9, dilation=3)
    om = oracle.DeformConv2D(8, groups=8, kernel_size=(7, 7), padding=9, dilation=3)
    om.load_state_dict(m.state_dict())
    x = torch.randn(2, 8, 20, 17)
    with torch.no_grad():
        ref = om(x)
        got = m.to(DEV)(x.to(DEV))
    assert rel_err(got, ref) < TOL


# ----------------------------------------------------------------------------- K4: integer planes, bit exact
@pytest.mark.parametrize("scale", [0.5, 3.0, 40.0])
def test_sample_indices3d_bit_exact(dl, oracle, scale):
    torch.manual_seed(7)
    D, H, W = 5, 6, 7
    off = torch.randn(2, 2 * 81, D, H, W) * scale
    off[0, :, 0, 0, 0] = torch.round(off[0, :, 0, 0, 0])  # exactly integral positions
    off[0, :3, 1, 1, 1] = torch.tensor([-1.0, -1.0, -1.0])  # p == -1 boundary on tap 0 at voxel (1,1,1): p = 1-1-1 = -1
    lo_ref, m_ref = oracle.sample_indices3d(off, (D, H, W), 3, 1, 1, 1, deformable_groups=2)
    lo, m = dl.ops.deform_conv3d_sample_indices(off.to(DEV), (D, H, W), 3, 1, 1, 1, deformable_group=2)
    assert torch.equal(m.cpu(), m_ref)
    valid = (m_ref & 1).bool()
    assert valid.any() and (~valid).any()
    assert torch.equal(lo.cpu()[valid], lo_ref[valid])


def test_sample_indices2d_bit_exact(dl, oracle):
    torch.manual_seed(8)
    H, W = 9, 8
    off = torch.randn(2, 98, H, W) * 4
    lo_ref, m_ref = oracle.sample_indices2d(off, (H, W), 7, 1, 9, 3)
    lo, m = dl.ops.deform_conv2d_sample_indices(off.to(DEV), (H, W), 7, 1, 9, 3)
    assert torch.equal(m.cpu(), m_ref)
    valid = (m_ref & 1).bool()
    assert torch.equal(lo.cpu()[valid], lo_ref[valid])


# ----------------------------------------------------------------------------- blocks vs oracle
def _scale_offset_nets(m, scale):
    with torch.no_grad():
        for name, mod in m.named_modules():
            if name.endswith("offset_net"):
                mod.weight.mul_(scale); mod.bias.mul_(scale)


@pytest.mark.parametrize("C,H,W,scale", [(8, 14, 12, 1.0), (64, 20, 24, 1.0), (96, 12, 10, 4.0), (12, 9, 21, 8.0)])
def test_block2d_vs_oracle(dl, oracle, C, H, W, scale, math):
    torch.manual_seed(9)
    ref_m = oracle.deformable_LKA_Attention(C).eval()
    _scale_offset_nets(ref_m, scale)
    m = dl.deformable_LKA_Attention(C)
    m.load_state_dict(ref_m.state_dict())
    x = torch.randn(2, C, H, W)
    with torch.no_grad():
        ref = ref_m(x)
        got = m.to(DEV)(x.to(DEV))
        ref_l = ref_m.spatial_gating_unit(x)
        got_l = m.spatial_gating_unit(x.to(DEV))
    assert rel_err(got, ref) < TOL
    assert rel_err(got_l, ref_l) < TOL


@pytest.mark.parametrize("C,dims,std,br", [(8, (6, 5, 4), 0.0, 0.0), (32, (8, 8, 8), 0.05, 1.0), (64, (4, 6, 5), 0.05, 1.0),
                                           (96, (5, 9, 8), 0.1, 6.0), (128, (4, 4, 4), 0.05, 1.0), (256, (4, 4, 4), 0.05, 1.0)])
def test_block3d_vs_oracle(dl, oracle, C, dims, std, br, math):
    torch.manual_seed(10)
    H, W, D = dims
    B = 2
    ref_m = oracle.LKA_Attention3d_deform(C).eval()
    if std > 0:
        oracle.randomize_offsets_(ref_m, std=std, bias_range=br)
    m = dl.LKA_Attention3d_deform(C)
    m.load_state_dict(ref_m.state_dict())
    m = m.to(DEV)
    x = torch.randn(B, H * W * D, C)
    with torch.no_grad():
        ref = ref_m(x, B, C, H, W, D)
        got = m(x.to(DEV), B, C, H, W, D)
        xv = torch.randn(B, C, H, W, D)
        ref_l = ref_m.spatial_gating_unit(xv)
        got_l = m.spatial_gating_unit(xv.to(DEV))
    assert got.shape == ref.shape
    assert rel_err(got, ref) < TOL
    assert rel_err(got_l, ref_l) < TOL


# ACDC variant of the block (row N4): stencil shapes per channel count, acdc/transformerblock.py:214-236.  The volumes are
# ragged against every tile shape (lattice tiles, bricks) and thinner than the (5,7,7)-dil-3 reach along the first axis.
@pytest.mark.parametrize("C,dims", [(32, (4, 13, 11)), (64, (6, 9, 10)), (128, (3, 8, 9)), (256, (2, 5, 5)), (32, (16, 20, 20))])
def test_block3d_acdc_vs_oracle(dl, oracle, C, dims, math):
    from deformablelka_b200 import acdc
    torch.manual_seed(12)
    H, W, D = dims
    B = 2
    ref_m = oracle.LKA_Attention3d_deform_ACDC(C).eval()
    oracle.randomize_offsets_(ref_m, std=0.05, bias_range=1.0)
    m = acdc.LKA_Attention3d_deform(C)
    m.load_state_dict(ref_m.state_dict())
    m = m.to(DEV)
    x = torch.randn(B, H * W * D, C)
    with torch.no_grad():
        ref = ref_m(x, B, C, H, W, D)
        got = m(x.to(DEV), B, C, H, W, D)
        xv = torch.randn(B, C, H, W, D)
        ref_l = ref_m.spatial_gating_unit(xv)
        got_l = m.spatial_gating_unit(xv.to(DEV))
    assert got.shape == ref.shape
    assert rel_err(got, ref) < TOL
    assert rel_err(got_l, ref_l) < TOL


def test_transformer3d_whole_block_acdc_vs_oracle(dl, oracle, math):
    """The ACDC network's transformer block (acdc/transformerblock.py:146-207): attention half with the ACDC stencil shapes,
    UnetResBlock and conv8, one library call, against the oracle's block restatement with the ACDC attention swapped in."""
    from deformablelka_b200 import acdc
    torch.manual_seed(13)
    C, (H, W, D) = 128, (3, 6, 5)
    N = H * W * D
    m = acdc.TransformerBlock_3D_single_deform_LKA(N, C, C, 4, pos_embed=True).eval()
    assert isinstance(m.epa_block, acdc.LKA_Attention3d_deform)
    oracle.randomize_offsets_(m)
    with torch.no_grad():
        m.gamma.uniform_(0.2, 1.0)
        m.pos_embed.normal_(0, 0.5)
        for bn in (m.conv51.norm1, m.conv51.norm2):
            bn.running_mean.normal_(0, 0.2); bn.running_var.uniform_(0.5, 1.5); bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.2)
    ref_attn = oracle.LKA_Attention3d_deform_ACDC(C).eval()
    ref_attn.load_state_dict(m.epa_block.state_dict())
    ref_res = oracle.UnetResBlock3D(C).eval()
    ref_res.load_state_dict(m.conv51.state_dict())
    x = torch.randn(2, C, H, W, D)
    with torch.no_grad():
        ref = oracle.transformer3d_block(m.norm, m.gamma, ref_attn, m.pos_embed, ref_res, m.conv8[1], x)
        got = m.to(DEV)(x.to(DEV))
    assert got.shape == ref.shape
    assert rel_err(got, ref) < TOL


def test_block3d_unsupported_stencil_is_loud(dl):
    """A stencil shape the library has no kernel for (k along axis 2 != axis 3) is refused, not approximated."""
    from deformablelka_b200 import acdc
    m = acdc.LKA_Attention3d_deform(32).to(DEV)
    m.spatial_gating_unit.dw_geom = ((5, 5, 5), (1, 1, 1), (5, 7, 5), (3, 3, 3))
    x = torch.randn(1, 4 * 4 * 4, 32, device=DEV)
    with pytest.raises(RuntimeError), torch.no_grad():
        m(x, 1, 32, 4, 4, 4)


# ----------------------------------------------------------------------------- properties at larger size
def test_block3d_batch_shard_consistency_and_identity(dl, oracle, math):
    """Size-independent properties on a mid-size volume: (a) B=2 equals two B=1 runs (the data-parallel
    split of SURVEY.md 8e is exact); (b) zero-initialised conv_offset == the same block with a plain
    Conv3d run by torch on the GPU (K3 at block level)."""
    torch.manual_seed(11)
    C, H, W, D = 32, 24, 20, 28
    m = dl.LKA_Attention3d_deform(C).to(DEV)
    x = torch.randn(2, H * W * D, C, device=DEV)
    with torch.no_grad():
        y = m(x, 2, C, H, W, D)
        y0 = m(x[:1].contiguous(), 1, C, H, W, D)
        y1 = m(x[1:].contiguous(), 1, C, H, W, D)
        assert torch.equal(y, torch.cat([y0, y1]))
        sg = m.spatial_gating_unit
        xx = x.permute(0, 2, 1).reshape(2, C, H, W, D)
        prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
        torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
        try:
            t = F.gelu(m.proj_1(xx))
            a = F.conv3d(sg.conv_spatial(sg.conv0(t)), sg.deform_conv.weight, sg.deform_conv.bias, 1, 1)
            ref = (m.proj_2(t * sg.conv1(a)) + xx).reshape(2, C, -1).permute(0, 2, 1)
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
    assert rel_err(y, ref) < TOL


def test_streams_and_noncontiguous_inputs(dl):
    torch.manual_seed(12)
    m = dl.deformable_LKA_Attention(16).to(DEV).requires_grad_(False)
    x = torch.randn(2, 16, 10, 12, device=DEV)
    y = m(x)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        y2 = m(x)
    s.synchronize()
    assert torch.equal(y, y2)
    xt = x.permute(0, 1, 3, 2)  # non-contiguous view: the host makes it contiguous like nn.Conv2d would accept
    assert torch.equal(m(xt), m(xt.contiguous()))
    assert dl.launch_count() > 0


def test_forward_host_pipeline_matches_device_forward(dl, oracle, math):
    """The host-buffer entry (per-sample H2D / compute / D2H pipeline over three streams) returns the same bits."""
    torch.manual_seed(13)
    C, H, W, D, B = 32, 12, 10, 9, 3
    m = dl.LKA_Attention3d_deform(C)
    oracle.randomize_offsets_(m)
    m = m.to(DEV)
    xh = torch.randn(B, H * W * D, C).pin_memory()
    with torch.no_grad():
        ref = m(xh.to(DEV), B, C, H, W, D).cpu()
        got = m.forward_host(xh, B, C, H, W, D)
        got2 = m.forward_host(xh, B, C, H, W, D, y_host=torch.empty_like(xh))  # pageable output buffer
    assert got.device.type == "cpu" and torch.equal(got, ref) and torch.equal(got2, ref)


@pytest.mark.parametrize("math_mode", ["bf16x3"])
def test_headline_shape_properties(dl, math_mode, monkeypatch):
    """BASELINE.json headline size (2,96,64,128,128): the oracle cannot run here in seconds, so check the
    size-independent properties on the GPU: (a) zero-initialised conv_offset block == the same block composed from
    torch's own CUDA convolutions in fp32 (K3 identity, exercises 64-bit addressing: 201M elements per tensor,
    row index x 96 > 2^31 bytes); (b) batch-shard exactness (sample 1 alone == sample 1 inside the batch);
    (c) with random offsets: permuting the batch permutes the output (no cross-sample coupling)."""
    monkeypatch.setenv("DLKA_MATH", math_mode)
    torch.manual_seed(14)
    B, C, H, W, D = 2, 96, 64, 128, 128
    m = dl.LKA_Attention3d_deform(C).to(DEV)
    x = torch.randn(B, H * W * D, C, device=DEV)
    with torch.no_grad():
        y = m(x, B, C, H, W, D)
        assert torch.isfinite(y).all()
        # (a) K3 at full size, one sample (keeps torch's workspace small)
        sg = m.spatial_gating_unit
        prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
        torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
        try:
            xx = x[1:].permute(0, 2, 1).reshape(1, C, H, W, D)
            t = F.gelu(m.proj_1(xx))
            a = F.conv3d(sg.conv_spatial(sg.conv0(t)), sg.deform_conv.weight, sg.deform_conv.bias, 1, 1)
            ref = (m.proj_2(t * sg.conv1(a)) + xx).reshape(1, C, -1).permute(0, 2, 1)
            del t, a
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
        err = ((y[1:] - ref).abs().max() / ref.abs().max()).item()
        assert err < TOL, err
        del ref
        # (b) + (c) with non-trivial offsets
        co = sg.deform_conv.conv_offset
        co.weight.normal_(0, 0.05); co.bias.uniform_(-1, 1)
        y2 = m(x, B, C, H, W, D)
        y1 = m(x[1:].contiguous(), 1, C, H, W, D)
        assert torch.equal(y2[1:], y1)
        yp = m(x.flip(0).contiguous(), B, C, H, W, D)
        assert torch.equal(yp.flip(0), y2)
        assert (y2 - y).abs().max() > 1e-3


def test_host_pipe_streaming_matches_device_forward(dl, oracle):
    """Streaming host pipeline (depth 2, 5 steps in flight order) returns the same bits for every step."""
    torch.manual_seed(15)
    C, H, W, D, B = 32, 8, 10, 6, 2
    m = dl.LKA_Attention3d_deform(C)
    oracle.randomize_offsets_(m)
    m = m.to(DEV)
    xs = [torch.randn(B, H * W * D, C).pin_memory() for _ in range(5)]
    ys = [torch.empty(B, H * W * D, C).pin_memory() for _ in range(5)]
    pipe = m.host_pipe(depth=2)
    with torch.no_grad():
        for x, y in zip(xs, ys):
            m.submit_host(pipe, x, y, B, C, H, W, D)
        pipe.wait()
        for x, y in zip(xs, ys):
            assert torch.equal(y, m(x.to(DEV), B, C, H, W, D).cpu())


# ----------------------------------------------------------------------------- row N1: enclosing blocks
@pytest.mark.parametrize("C,H,W,scale", [(16, 10, 12, 1.0), (96, 14, 14, 1.0), (64, 9, 20, 6.0)])
def test_lka_block2d_vs_oracle(dl, oracle, C, H, W, scale, math):
    torch.manual_seed(16)
    ref_m = oracle.deformableLKABlock(C).eval()
    _scale_offset_nets(ref_m, scale)
    with torch.no_grad():
        ref_m.layer_scale_1.uniform_(0.2, 1.0); ref_m.layer_scale_2.uniform_(0.2, 1.0)  # make both branches count
        ref_m.norm1.weight.uniform_(0.5, 1.5); ref_m.norm1.bias.normal_(0, 0.1)
    m = dl.deformableLKABlock(C)
    assert sorted(m.state_dict().keys()) == sorted(ref_m.state_dict().keys())
    m.load_state_dict(ref_m.state_dict())
    x = torch.randn(2, H * W, C)
    with torch.no_grad():
        ref = ref_m(x, H, W)
        got = m.to(DEV)(x.to(DEV), H, W)
    assert got.shape == ref.shape
    assert rel_err(got, ref) < TOL


# ----------------------------------------------------------------------------- rest of row N3: the 2D decoder stage
@pytest.mark.parametrize("dim,H,W,scale", [(96, 7, 7, 2), (192, 5, 6, 2), (768, 3, 2, 2), (96, 6, 5, 4), (32, 4, 4, 4)])
def test_patch_expand2d_vs_oracle(dl, oracle, dim, H, W, scale, math):
    torch.manual_seed(20)
    ref_m = (oracle.PatchExpand((H, W), dim) if scale == 2 else oracle.FinalPatchExpand_X4((H, W), dim)).eval()
    with torch.no_grad():
        ref_m.norm.weight.uniform_(0.5, 1.5); ref_m.norm.bias.normal_(0, 0.2)
    m = dl.PatchExpand((H, W), dim) if scale == 2 else dl.FinalPatchExpand_X4((H, W), dim)
    assert sorted(m.state_dict().keys()) == sorted(ref_m.state_dict().keys())
    m.load_state_dict(ref_m.state_dict())
    x = torch.randn(2, H * W, dim)
    with torch.no_grad():
        ref = ref_m(x)
        got = m.to(DEV)(x.to(DEV))
    assert got.shape == ref.shape
    assert rel_err(got, ref) < TOL
    with pytest.raises(AssertionError):
        m(torch.randn(2, H * W + 1, dim, device=DEV))   # "input feature has wrong size"


@pytest.mark.parametrize("dim,H,W,last", [(96, 6, 5, False), (32, 4, 6, True), (192, 3, 4, False)])
def test_decoder_layer2d_vs_oracle(dl, oracle, dim, H, W, last, math):
    torch.manual_seed(21)
    chans = [dim] * 5
    ref_m = oracle.MyDecoderLayer((H, W), chans, 1, "mix_skip", n_class=9, is_last=last).eval()
    _scale_offset_nets(ref_m, 1.0)
    with torch.no_grad():
        for blk in (ref_m.layer_lka_1, ref_m.layer_lka_2):
            blk.layer_scale_1.uniform_(0.2, 1.0); blk.layer_scale_2.uniform_(0.2, 1.0)
        ref_m.x1_linear.bias.normal_(0, 0.2)
    m = dl.MyDecoderLayer((H, W), chans, 1, "mix_skip", n_class=9, is_last=last)
    assert sorted(m.state_dict().keys()) == sorted(ref_m.state_dict().keys())
    m.load_state_dict(ref_m.state_dict())
    m = m.to(DEV)
    x1, x2 = torch.randn(2, H * W, dim), torch.randn(2, H, W, dim)
    with torch.no_grad():
        ref = ref_m(x1, x2)
        got = m(x1.to(DEV), x2.to(DEV))
        ref0 = ref_m(x1)                 # no skip connection: only the patch expansion (:618-619)
        got0 = m(x1.to(DEV))
    assert got.shape == ref.shape and got0.shape == ref0.shape
    assert rel_err(got, ref) < TOL
    assert rel_err(got0, ref0) < TOL


@pytest.mark.parametrize("C,dims,pos", [(32, (6, 5, 8), True), (96, (4, 6, 5), False)])
def test_transformer3d_attention_half_vs_oracle(dl, oracle, C, dims, pos, math):
    torch.manual_seed(17)
    H, W, D = dims
    N = H * W * D
    m = dl.TransformerBlock_3D_single_deform_LKA(N, C, C, 4, pos_embed=pos)
    oracle.randomize_offsets_(m)
    with torch.no_grad():
        m.gamma.uniform_(0.2, 1.0)          # reference initialises 1e-6: make the branch visible
        if pos:
            m.pos_embed.normal_(0, 0.5)
    ref_attn = oracle.LKA_Attention3d_deform(C).eval()
    ref_attn.load_state_dict(m.epa_block.state_dict())
    x = torch.randn(2, C, H, W, D)
    with torch.no_grad():
        ref = oracle.transformer3d_attention_half(m.norm, m.gamma, ref_attn, m.pos_embed, x)
        md = m.to(DEV).eval()
        tok = x.to(DEV).reshape(2, C, N).permute(0, 2, 1).contiguous()
        got = md.attention_half(tok, 2, C, H, W, D)
        full = md(x.to(DEV))                 # whole block incl. the stock-PyTorch UnetResBlock tail (row N3)
    assert rel_err(got, ref) < TOL
    assert full.shape == x.shape and torch.isfinite(full).all()


@pytest.mark.parametrize("C,dims,pos", [(32, (6, 5, 8), True), (96, (4, 6, 5), False), (256, (4, 4, 4), True)])
def test_transformer3d_whole_block_vs_oracle(dl, oracle, C, dims, pos, math):
    """Rows N1 + N3: the whole TransformerBlock_3D_single_deform_LKA (attention half + UnetResBlock + conv8)."""
    torch.manual_seed(18)
    H, W, D = dims
    N = H * W * D
    m = dl.TransformerBlock_3D_single_deform_LKA(N, C, C, 4, pos_embed=pos)
    oracle.randomize_offsets_(m)
    with torch.no_grad():
        m.gamma.uniform_(0.2, 1.0)
        if pos:
            m.pos_embed.normal_(0, 0.5)
        for bn in (m.conv51.norm1, m.conv51.norm2):   # non-trivial running statistics
            bn.running_mean.normal_(0, 0.3); bn.running_var.uniform_(0.5, 2.0); bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.2)
    m.eval()
    ref_attn = oracle.LKA_Attention3d_deform(C).eval()
    ref_attn.load_state_dict(m.epa_block.state_dict())
    ref_res = oracle.UnetResBlock3D(C).eval()
    ref_res.load_state_dict(m.conv51.state_dict())
    x = torch.randn(2, C, H, W, D)
    with torch.no_grad():
        ref = oracle.transformer3d_block(m.norm, m.gamma, ref_attn, m.pos_embed, ref_res, m.conv8[1], x)
        md = m.to(DEV)
        got = md(x.to(DEV))
        got_tail_torch = md.forward_reference_tail(x.to(DEV))
    assert got.shape == ref.shape
    assert rel_err(got, ref) < TOL
    assert rel_err(got_tail_torch, ref) < TOL
    with pytest.raises(RuntimeError, match="eval"):   # the fused entry is inference-only ...
        md.train().forward_tokens(x.to(DEV).reshape(2, C, N).permute(0, 2, 1).contiguous(), 2, C, H, W, D)
    md.eval()


# ----------------------------------------------------------------------------- row N2: backward of the 3D deformable conv
@pytest.mark.parametrize("B,C,Co,dims,k,stride,pad,dil,scale", [
    (2, 8, 12, (4, 5, 6), 3, 1, 1, 1, 1.5),        # offsets beyond the volume: validity / corner masks in the gradients
    (1, 32, 32, (4, 4, 4), 3, 1, 1, 1, 0.3),
    (2, 16, 8, (7, 6, 5), 3, 2, 1, 1, 0.7),        # stride 2
    (1, 8, 8, (6, 7, 8), 3, 1, 2, 2, 0.7),         # dilation 2
    (1, 8, 16, (5, 5, 5), (1, 3, 3), 1, (0, 1, 1), 1, 0.5),
    (1, 8, 8, (24, 20, 20), 3, 1, 1, 1, 0.5),      # 9600 rows: two streamed chunks (accumulating grad_weight GEMM, zero padding)
])
def test_deform_conv3d_backward_vs_autograd_oracle(dl, oracle, B, C, Co, dims, k, stride, pad, dil, scale, math):
    from torch.nn.modules.utils import _triple
    torch.manual_seed(30)
    D, H, W = dims
    kd, kh, kw = _triple(k)
    st, pd, dl_ = _triple(stride), _triple(pad), _triple(dil)
    ext = [(n + 2 * p - (d * (kk - 1) + 1)) // s + 1 for n, p, d, kk, s in zip(dims, pd, dl_, (kd, kh, kw), st)]
    x = torch.randn(B, C, D, H, W, requires_grad=True)
    w = (torch.randn(Co, C, kd, kh, kw) * 0.2).requires_grad_()
    b = torch.randn(Co, requires_grad=True)
    off = (torch.randn(B, 3 * kd * kh * kw, *ext) * scale).requires_grad_()
    gout = torch.randn(B, Co, *ext)
    ref = oracle.deform_conv3d_autograd(x, off, w, b, st, pd, dl_)
    ref.backward(gout)
    gi, go, gw, gb = dl.ops.deform_conv3d_backward(x.detach().to(DEV), w.detach().to(DEV), b.detach().to(DEV), off.detach().to(DEV),
                                                   gout.to(DEV), (kd, kh, kw), st, pd, dl_, 1, 1)
    for name, got, want in (("grad_input", gi, x.grad), ("grad_offset", go, off.grad), ("grad_weight", gw, w.grad),
                            ("grad_bias", gb, b.grad)):
        assert got.shape == want.shape, name
        assert rel_err(got, want) < TOL, name


def test_deform_conv_pack3d_autograd_end_to_end(dl, oracle, math):
    """loss.backward() through the drop-in module (DeformConvFunction.backward, deform_conv_func.py:38-56): gradients of
    every parameter and of the input against autograd through the oracle restatement."""
    torch.manual_seed(31)
    C, dims = 16, (5, 6, 4)
    m = dl.DeformConvPack(C, C, (3, 3, 3), 1, 1)
    oracle.randomize_offsets_(m, std=0.1, bias_range=1.0)
    mo = oracle.DeformConvPack3D(C, C, (3, 3, 3), 1, 1)
    mo.load_state_dict(m.state_dict())
    x = torch.randn(2, C, *dims)
    xo = x.clone().requires_grad_()
    off = mo.conv_offset(xo)
    oracle.deform_conv3d_autograd(xo, off, mo.weight, mo.bias).square().sum().backward()
    m = m.to(DEV)
    xg = x.to(DEV).requires_grad_()
    m(xg).square().sum().backward()
    assert rel_err(xg.grad, xo.grad) < TOL
    for (n, p), (_, po) in zip(m.named_parameters(), mo.named_parameters()):
        assert rel_err(p.grad, po.grad) < TOL, n


@pytest.mark.parametrize("C,Co,group,dgrp,dims,scale", [
    (16, 8, 2, 1, (4, 5, 6), 0.7),      # weight groups only
    (16, 16, 1, 2, (5, 4, 6), 1.5),     # deformable groups only (each half of the channels has its own sampling grid)
    (32, 24, 2, 4, (4, 4, 5), 0.7),     # both, different counts
    (16, 16, 4, 4, (3, 4, 4), 0.5),     # 4 channels per deformable group: one float4 per group
])
def test_deform_conv3d_backward_groups_vs_autograd_oracle(dl, oracle, C, Co, group, dgrp, dims, scale, math):
    """group / deformable_group != 1 (deform_conv_cuda.cu:160-166, 204-270) against autograd through the oracle, whose FORWARD with
    groups is first checked against the forward oracle (pinned to the compiled reference in tests/test_ref_d3d_gpu.py)."""
    torch.manual_seed(32)
    D, H, W = dims
    x = torch.randn(2, C, D, H, W, requires_grad=True)
    w = (torch.randn(Co, C // group, 3, 3, 3) * 0.2).requires_grad_()
    b = torch.randn(Co, requires_grad=True)
    off = (torch.randn(2, dgrp * 81, D, H, W) * scale).requires_grad_()
    gout = torch.randn(2, Co, D, H, W)
    ref = oracle.deform_conv3d_autograd(x, off, w, b, (1, 1, 1), (1, 1, 1), (1, 1, 1), group, dgrp)
    with torch.no_grad():
        fwd = oracle.deform_conv3d(x, off, w, b, 1, 1, 1, group, dgrp)
    assert rel_err(ref.detach(), fwd) < 1e-5
    ref.backward(gout)
    gi, go, gw, gb = dl.ops.deform_conv3d_backward(x.detach().to(DEV), w.detach().to(DEV), b.detach().to(DEV), off.detach().to(DEV),
                                                   gout.to(DEV), 3, 1, 1, 1, group, dgrp)
    for name, got, want in (("grad_input", gi, x.grad), ("grad_offset", go, off.grad), ("grad_weight", gw, w.grad),
                            ("grad_bias", gb, b.grad)):
        assert got.shape == want.shape, name
        assert rel_err(got, want) < TOL, name


def test_deform_conv3d_backward_refuses_indivisible_groups(dl):
    x = torch.randn(1, 8, 3, 3, 3, device=DEV)
    w = torch.randn(9, 4, 3, 3, 3, device=DEV)   # 9 output channels in 2 groups
    with pytest.raises(RuntimeError):
        dl.ops.deform_conv3d_backward(x, w, torch.zeros(9, device=DEV), torch.zeros(1, 81, 3, 3, 3, device=DEV),
                                      torch.zeros(1, 9, 3, 3, 3, device=DEV), 3, 1, 1, 1, 2, 1)


# ----------------------------------------------------------------------------- 1x1 projections at large M (generic kernel by
# for M < 37888 rows, the persistent streaming kernel dense_stream.cu above that: both sizes are covered)
@pytest.mark.parametrize("M,K,N,bias,add", [(20000, 96, 96, True, True), (19001, 64, 128, True, False), (40000, 32, 9, False, True),
                                             (18944, 96, 81, True, False), (300001, 96, 96, True, True), (262144, 64, 64, True, False),
                                             (150000, 32, 40, False, True), (37888, 96, 81, True, False)])
def test_linear_tokens_large_m_vs_torch(dl, M, K, N, bias, add, math):
    """Many 128-row tiles per SM: ragged last tile, N not a multiple of 16, bias / residual epilogues."""
    torch.manual_seed(40)
    x = torch.randn(M, K)
    w = torch.randn(N, K) * 0.2
    b = torch.randn(N) if bias else None
    e = torch.randn(M, N) if add else None
    ref = torch.nn.functional.linear(x.double(), w.double(), b.double() if bias else None)
    if add:
        ref = ref + e.double()
    got = dl.ops.linear_tokens_forward(x.to(DEV), w.to(DEV), b.to(DEV) if bias else None, add=e.to(DEV) if add else None)
    assert got.shape == (M, N)
    assert rel_err(got, ref.float()) < TOL


def test_attention2d_medium_vs_oracle(dl, oracle, math):
    """2 x 100 x 100 pixels = 20000 rows through proj_1 + GELU, conv1 * u and proj_2 + x."""
    torch.manual_seed(41)
    C, H, W = 32, 100, 100
    ref_m = oracle.deformable_LKA_Attention(C).eval()
    _scale_offset_nets(ref_m, 1.0)
    m = dl.deformable_LKA_Attention(C)
    m.load_state_dict(ref_m.state_dict())
    x = torch.randn(2, C, H, W)
    with torch.no_grad():
        ref = ref_m(x)
        got = m.to(DEV)(x.to(DEV))
    assert rel_err(got, ref) < TOL


def test_numa_placed_pinned_buffers_and_pipe_slot_lifetime(dl, oracle):
    """ops.pinned_empty (dlka_host_alloc: mbind + cudaHostRegister) returns page-locked memory that the host pipe can stream
    from / to; steps with different batch sizes share a slot (the slot's "inputs consumed" event is per step, not per batch
    index); the pipe survives dropping every caller-side reference to the host tensors right after submit."""
    torch.manual_seed(42)
    C, H, W, D = 32, 6, 8, 5
    m = dl.LKA_Attention3d_deform(C)
    oracle.randomize_offsets_(m)
    m = m.to(DEV)
    assert dl.ops.bind_host_thread(DEV) >= -1
    pipe = m.host_pipe(depth=2)
    outs, refs = [], []
    with torch.no_grad():
        for step, B in enumerate((3, 1, 2, 3, 1)):
            x = dl.ops.pinned_empty((B, H * W * D, C), DEV)
            assert x.is_pinned()
            x.normal_()
            y = dl.ops.pinned_empty((B, H * W * D, C), DEV)
            refs.append(m(x.to(DEV), B, C, H, W, D).cpu())
            m.submit_host(pipe, x, y, B, C, H, W, D)
            outs.append(y)
            del x, y
        pipe.wait()
    for got, ref in zip(outs, refs):
        assert torch.equal(got, ref)

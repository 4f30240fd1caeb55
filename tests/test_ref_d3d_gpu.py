"""Pin the oracle AND the product against the reference's own D3D extension, compiled for sm_100a by
oracle/build_ref.py into oracle/_ref/ (two-token torch-2 patch, see that file).  D3D is CUDA-only, so
this runs on the GPU box only; it is skipped when oracle/_ref was not built."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def d3d():
    from oracle import build_ref
    mod = build_ref.load_d3d()
    if mod is None:
        pytest.skip("oracle/_ref/D3D*.so not built")
    return mod


def _ref(d3d, x, w, b, off, k, s, p, d, g, dg):
    k = (k,) * 3 if isinstance(k, int) else k
    s = (s,) * 3 if isinstance(s, int) else s
    p = (p,) * 3 if isinstance(p, int) else p
    d = (d,) * 3 if isinstance(d, int) else d
    return d3d.deform_conv_forward(x, w, b, off, *k, *s, *p, *d, g, dg, 64)


def rel_err(a, b):
    a = a.float().cpu(); b = b.float().cpu()
    return ((a - b).abs().max() / b.abs().max()).item()


@pytest.mark.parametrize("C,Co,g,dg,k,s,p,d,scale", [
    (8, 8, 1, 1, 3, 1, 1, 1, 1.0),
    (16, 12, 1, 1, 3, 1, 1, 1, 8.0),
    (16, 16, 2, 2, (3, 2, 3), (1, 2, 1), (1, 0, 2), (1, 2, 1), 2.0),
])
def test_oracle_matches_compiled_reference(d3d, oracle, C, Co, g, dg, k, s, p, d, scale):
    torch.manual_seed(0)
    B, D, H, W = 2, 6, 7, 9
    kd, kh, kw = oracle._triple(k); sd, sh, sw = oracle._triple(s); pd, ph, pw = oracle._triple(p); dd, dh, dw = oracle._triple(d)
    Do, Ho, Wo = oracle.out_extent(D, pd, dd, kd, sd), oracle.out_extent(H, ph, dh, kh, sh), oracle.out_extent(W, pw, dw, kw, sw)
    x = torch.randn(B, C, D, H, W); w = torch.randn(Co, C // g, kd, kh, kw) * 0.2; b = torch.randn(Co)
    off = torch.randn(B, dg * 3 * kd * kh * kw, Do, Ho, Wo) * scale
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = _ref(d3d, x.to(DEV), w.to(DEV), b.to(DEV), off.to(DEV), k, s, p, d, g, dg)
    ora = oracle.deform_conv3d(x, off, w, b, s, p, d, g, dg)
    assert rel_err(ora, ref) < 1e-5


def test_product_matches_compiled_reference_at_c3_shape(d3d):
    """BASELINE config 3: (2,64,32,64,64), k=3 -- the largest shape class the reference's int32 indexing survives."""
    import deformablelka_b200 as dl
    torch.manual_seed(1)
    B, C, D, H, W = 2, 64, 32, 64, 64
    x = torch.randn(B, C, D, H, W, device=DEV); w = torch.randn(C, C, 3, 3, 3, device=DEV) * 0.05
    b = torch.randn(C, device=DEV); off = torch.randn(B, 81, D, H, W, device=DEV)
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = _ref(d3d, x, w, b, off, 3, 1, 1, 1, 1, 1)
    for math in ("fp32", "bf16x3"):
        got = dl.ops.deform_conv3d_forward(x, w, b, off, 3, 1, 1, 1, 1, 1, 64, math=math)
        assert rel_err(got, ref) < 1e-3, math


@pytest.mark.parametrize("C,Co,dims,scale", [(8, 12, (6, 7, 9), 1.5), (32, 32, (8, 8, 8), 0.5), (16, 16, (24, 20, 20), 0.5)])
def test_backward_matches_compiled_reference(d3d, oracle, C, Co, dims, scale):
    """Row N2 against the reference's own D3D.deform_conv_backward (deform_conv_cuda.cu:128-285) at the block's configuration
    k = 3, stride 1, pad (1,1,1), dilation 1: with equal pads on every axis the reference's pad_h/pad_w index defect
    (deform_im2col_cuda.cuh:448) has no effect, so its four gradients are the exact target; the autograd oracle is checked
    against it on the way (pins the backward oracle to the real reference)."""
    import deformablelka_b200 as dl
    torch.manual_seed(2)
    B = 2
    D, H, W = dims
    x = torch.randn(B, C, D, H, W); w = torch.randn(Co, C, 3, 3, 3) * 0.2; b = torch.randn(Co)
    off = torch.randn(B, 81, D, H, W) * scale
    gout = torch.randn(B, Co, D, H, W)
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = d3d.deform_conv_backward(x.to(DEV), w.to(DEV), b.to(DEV), off.to(DEV), gout.to(DEV), 3, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 64)
    names = ("grad_input", "grad_offset", "grad_weight", "grad_bias")
    if x.numel() < 200000:   # the pure-torch autograd oracle is for small shapes
        xo, wo, bo, oo = (t.clone().requires_grad_() for t in (x, w, b, off))
        oracle.deform_conv3d_autograd(xo, oo, wo, bo).backward(gout)
        for n, o_, r in zip(names, (xo.grad, oo.grad, wo.grad, bo.grad), ref):
            assert rel_err(o_, r) < 2e-5, "oracle " + n
    for math in ("fp32", "bf16x3"):
        got = dl.ops.deform_conv3d_backward(x.to(DEV), w.to(DEV), b.to(DEV), off.to(DEV), gout.to(DEV), 3, 1, 1, 1, 1, 1, 64, math=math)
        for n, g_, r in zip(names, got, ref):
            assert rel_err(g_, r) < 1e-3, f"{math} {n}"


@pytest.mark.parametrize("C,Co,g,dg,dims,scale", [(16, 8, 2, 1, (5, 6, 7), 0.7), (16, 16, 1, 2, (6, 5, 7), 1.5), (32, 24, 2, 4, (4, 6, 5), 0.7)])
def test_backward_groups_match_compiled_reference(d3d, oracle, C, Co, g, dg, dims, scale):
    """group / deformable_group != 1 in the backward (deform_conv_cuda.cu:160-166, 204-270), against the reference's own compiled
    D3D.deform_conv_backward; same equal-pad argument as above.  The grouped autograd oracle is pinned on the way."""
    import deformablelka_b200 as dl
    torch.manual_seed(3)
    B = 2
    D, H, W = dims
    x = torch.randn(B, C, D, H, W); w = torch.randn(Co, C // g, 3, 3, 3) * 0.2; b = torch.randn(Co)
    off = torch.randn(B, dg * 81, D, H, W) * scale
    gout = torch.randn(B, Co, D, H, W)
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = d3d.deform_conv_backward(x.to(DEV), w.to(DEV), b.to(DEV), off.to(DEV), gout.to(DEV), 3, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, g, dg, 64)
    names = ("grad_input", "grad_offset", "grad_weight", "grad_bias")
    xo, wo, bo, oo = (t.clone().requires_grad_() for t in (x, w, b, off))
    oracle.deform_conv3d_autograd(xo, oo, wo, bo, (1, 1, 1), (1, 1, 1), (1, 1, 1), g, dg).backward(gout)
    for n, o_, r in zip(names, (xo.grad, oo.grad, wo.grad, bo.grad), ref):
        assert rel_err(o_, r) < 2e-5, "oracle " + n
    for math in ("fp32", "bf16x3"):
        got = dl.ops.deform_conv3d_backward(x.to(DEV), w.to(DEV), b.to(DEV), off.to(DEV), gout.to(DEV), 3, 1, 1, 1, g, dg, 64, math=math)
        for n, g_, r in zip(names, got, ref):
            assert got[0].shape == ref[0].shape
            assert rel_err(g_, r) < 1e-3, f"{math} {n}"

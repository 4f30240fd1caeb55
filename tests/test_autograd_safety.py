"""A fused module must never hand back a tensor that is silently cut off from the autograd graph.

CPU-collectable part: the routing logic.  The 3D modules, the 3D transformer block and the patch expansions take a
differentiable composition when a gradient can be asked for; here the library's operator (CUDA-only) is replaced by the
oracle's CPU restatement so that the composition itself can run, and the gradients are compared with autograd through the
oracle's modules.  The 2D fused entries (no 2D backward in the library) must raise.  The GPU part backpropagates through the
real library operator.
"""
from unittest import mock

import pytest
import torch

TOL = 1e-3


def rel_err(got, ref):
    got = got.detach().float().cpu(); ref = ref.detach().float().cpu()
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


@pytest.fixture
def cpu_operator(oracle, monkeypatch):
    """Stand-in for the CUDA-only 3D operator so that the ROUTING can be tested without a GPU (test infrastructure only)."""
    import deformablelka_b200 as dl

    def fwd(input, weight, bias, offset, kernel_size, stride, padding, dilation, group, deformable_group, im2col_step=64, math=None):
        return oracle.deform_conv3d(input, offset, weight, bias, stride, padding, dilation, group, deformable_group)

    def bwd(input, weight, bias, offset, grad_output, kernel_size, stride, padding, dilation, group, deformable_group,
            im2col_step=64, math=None):
        with torch.enable_grad():
            x, w, b, o = (t.detach().clone().requires_grad_() for t in (input, weight, bias, offset))
            oracle.deform_conv3d_autograd(x, o, w, b, tuple(stride), tuple(padding), tuple(dilation)).backward(grad_output)
        return x.grad, o.grad, w.grad, b.grad

    monkeypatch.setattr(dl.ops, "deform_conv3d_forward", fwd)
    monkeypatch.setattr(dl.ops, "deform_conv3d_backward", bwd)

    def fused(*a, **k):
        raise AssertionError("the fused (non-differentiable) entry was taken although a gradient is required")

    for name in ("lka3d_deform_forward", "lka_attention3d_deform_forward", "lka_transformer3d_block_forward",
                 "lka_transformer3d_prenorm_forward", "deform_conv_pack3d", "patch_expand2d_forward"):
        monkeypatch.setattr(dl.ops, name, fused)
    # the differentiable compositions refuse CPU tensors like every other entry: pretend the CPU tensors are CUDA tensors
    with mock.patch.object(torch.Tensor, "is_cuda", new_callable=mock.PropertyMock, return_value=True):
        yield dl


def _oracle_attention_grads(oracle, m, x, dims, acdc=False):
    C = x.shape[-1]
    om = (oracle.LKA_Attention3d_deform_ACDC if acdc else oracle.LKA_Attention3d_deform)(C)
    om.load_state_dict(m.state_dict())
    sg = om.spatial_gating_unit
    xo = x.detach().clone().requires_grad_()
    B = x.shape[0]
    v = xo.permute(0, 2, 1).reshape(B, C, *dims)
    t = om.activation(om.proj_1(v))
    a = sg.conv_spatial(sg.conv0(t))
    dc = sg.deform_conv
    a = oracle.deform_conv3d_autograd(a, dc.conv_offset(a), dc.weight, dc.bias)
    y = (om.proj_2(t * sg.conv1(a)) + v).reshape(B, C, -1).permute(0, 2, 1)
    return om, xo, y


def test_attention3d_routes_to_differentiable_path_cpu(cpu_operator, oracle):
    dl = cpu_operator
    torch.manual_seed(0)
    C, dims = 8, (4, 5, 3)
    m = dl.LKA_Attention3d_deform(C)
    oracle.randomize_offsets_(m, std=0.1, bias_range=0.7)
    x = torch.randn(2, dims[0] * dims[1] * dims[2], C, requires_grad=True)
    y = m(x, 2, C, *dims)
    assert y.grad_fn is not None
    y.square().sum().backward()
    om, xo, yo = _oracle_attention_grads(oracle, m, x, dims)
    yo.square().sum().backward()
    assert rel_err(y, yo) < 1e-5
    assert rel_err(x.grad, xo.grad) < 1e-4
    for (n, p), (_, po) in zip(m.named_parameters(), om.named_parameters()):
        assert p.grad is not None, n
        assert rel_err(p.grad, po.grad) < 1e-4, n
    # parameters frozen + no_grad: the fused entry is what runs (here: the stub that raises AssertionError)
    with torch.no_grad(), pytest.raises(AssertionError, match="fused"):
        m(x, 2, C, *dims)


def test_transformer_block3d_train_and_eval_grad_cpu(cpu_operator, oracle):
    dl = cpu_operator
    torch.manual_seed(1)
    C, dims = 8, (3, 4, 3)
    N = dims[0] * dims[1] * dims[2]
    m = dl.TransformerBlock_3D_single_deform_LKA(N, C, C, 4, pos_embed=True)
    oracle.randomize_offsets_(m)
    with torch.no_grad():
        m.gamma.fill_(0.5)
    x = torch.randn(2, C, *dims)
    for mode in ("train", "eval"):
        getattr(m, mode)()
        m.zero_grad()
        y = m(x)                       # parameters require grad -> differentiable path in both modes
        assert y.shape == x.shape and y.grad_fn is not None
        y.square().sum().backward()
        missing = [n for n, p in m.named_parameters() if p.grad is None]
        assert not missing, missing
        assert m.epa_block.spatial_gating_unit.deform_conv.weight.grad.abs().max() > 0
    m.eval()
    with pytest.raises(RuntimeError, match="gradient is required"):
        m.forward_tokens(x.reshape(2, C, N).permute(0, 2, 1).contiguous(), 2, C, *dims)


def test_patch_expand_grad_path_matches_oracle_cpu(cpu_operator, oracle):
    dl = cpu_operator
    torch.manual_seed(2)
    for cls, ocls, scale in ((dl.PatchExpand, oracle.PatchExpand, 2), (dl.FinalPatchExpand_X4, oracle.FinalPatchExpand_X4, 4)):
        m, om = cls((3, 4), 16), ocls((3, 4), 16)
        om.load_state_dict(m.state_dict())
        x = torch.randn(2, 12, 16, requires_grad=True)
        y = m(x)
        assert y.grad_fn is not None and rel_err(y, om(x)) < 1e-6
        y.sum().backward()
        assert m.expand.weight.grad is not None


@pytest.fixture
def cpu_operator2d(monkeypatch):
    """Stand-in for the CUDA-only 2D operator (forward and backward) built on torchvision's CPU op: routing tests only."""
    import torchvision
    import deformablelka_b200 as dl

    def fwd(input, offset, weight, bias=None, stride=(1, 1), padding=(0, 0), dilation=(1, 1), mask=None, math=None):
        return torchvision.ops.deform_conv2d(input, offset, weight, bias, stride, padding, dilation, mask)

    def bwd(input, offset, weight, mask, grad_output, stride=(1, 1), padding=(0, 0), dilation=(1, 1), need_bias_grad=False):
        with torch.enable_grad():
            x, o, w = (t.detach().clone().requires_grad_() for t in (input, offset, weight))
            m = None if mask is None else mask.detach().clone().requires_grad_()
            b = torch.zeros(weight.shape[0], requires_grad=True) if need_bias_grad else None
            torchvision.ops.deform_conv2d(x, o, w, b, stride, padding, dilation, m).backward(grad_output)
        return x.grad, o.grad, w.grad, None if m is None else m.grad, None if b is None else b.grad

    monkeypatch.setattr(dl.ops, "deform_conv2d", fwd)
    monkeypatch.setattr(dl.ops, "deform_conv2d_backward", bwd)

    def fused(*a, **k):
        raise AssertionError("the fused (non-differentiable) entry was taken although a gradient is required")

    for name in ("deformable_lka2d_forward", "deformable_lka_attention2d_forward", "deformable_lka_block2d_forward",
                 "deform_conv_pack2d", "linear_tokens_forward", "patch_expand2d_forward"):
        monkeypatch.setattr(dl.ops, name, fused)
    with mock.patch.object(torch.Tensor, "is_cuda", new_callable=mock.PropertyMock, return_value=True):
        yield dl


def test_block2d_routes_to_differentiable_path_cpu(cpu_operator2d, oracle):
    """deformable_LKA_Attention / deformableLKABlock / MyDecoderLayer with grad required: composed from stock layers around the
    differentiable 2D operator; gradients equal autograd through the oracle's modules (torchvision's own autograd)."""
    dl = cpu_operator2d
    torch.manual_seed(4)
    C, H, W = 8, 7, 6
    m, om = dl.deformableLKABlock(C), oracle.deformableLKABlock(C)
    with torch.no_grad():
        m.layer_scale_1.uniform_(0.2, 1.0); m.layer_scale_2.uniform_(0.2, 1.0)
    om.load_state_dict(m.state_dict())
    x = torch.randn(2, H * W, C, requires_grad=True)
    xo = x.detach().clone().requires_grad_()
    y, yo = m(x, H, W), om(xo, H, W)
    assert y.grad_fn is not None and rel_err(y, yo) < 1e-5
    y.square().sum().backward(); yo.square().sum().backward()
    assert rel_err(x.grad, xo.grad) < 1e-4
    for (n, p), (_, po) in zip(m.named_parameters(), om.named_parameters()):
        assert p.grad is not None, n
        assert rel_err(p.grad, po.grad) < 1e-4, n
    # the decoder stage: every sub-module on its differentiable path
    d, od = dl.MyDecoderLayer((H, W), [C] * 5, 1, "mix_skip", is_last=True), oracle.MyDecoderLayer((H, W), [C] * 5, 1, "mix_skip", is_last=True)
    od.load_state_dict(d.state_dict())
    x1, x2 = torch.randn(1, H * W, C, requires_grad=True), torch.randn(1, H, W, C)
    out, oout = d(x1, x2), od(x1.detach(), x2)
    assert out.grad_fn is not None and rel_err(out, oout) < 1e-5
    out.sum().backward()
    assert all(p.grad is not None for p in d.parameters())
    # CPU tensors are refused either way (no CPU path)
    with mock.patch.object(torch.Tensor, "is_cuda", new_callable=mock.PropertyMock, return_value=False):
        with pytest.raises(RuntimeError, match="CPU"):
            dl.deformable_LKA_Attention(8)(torch.randn(1, 8, 5, 5))
        with pytest.raises(RuntimeError, match="CPU"):
            dl.PatchExpand((2, 2), 16)(torch.randn(1, 4, 16))


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("acdc", [False, True])
def test_attention3d_backward_on_gpu_vs_oracle(oracle, acdc):
    import deformablelka_b200 as dl
    torch.manual_seed(3)
    C, dims = 32, (5, 6, 4)
    m = (dl.acdc.LKA_Attention3d_deform if acdc else dl.LKA_Attention3d_deform)(C)
    oracle.randomize_offsets_(m, std=0.1, bias_range=0.7)
    x = torch.randn(2, dims[0] * dims[1] * dims[2], C)
    om, xo, yo = _oracle_attention_grads(oracle, m, x, dims, acdc)
    yo.square().sum().backward()
    m = m.to("cuda:0")
    xg = x.to("cuda:0").requires_grad_()
    with torch.backends.cudnn.flags(allow_tf32=False):
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            y = m(xg, 2, C, *dims)
            assert y.grad_fn is not None
            y.square().sum().backward()
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev
    with torch.no_grad():
        y_fused = m(xg.detach(), 2, C, *dims)          # the one-call inference path computes the same function
    assert rel_err(y, yo) < TOL and rel_err(y_fused, yo) < TOL
    assert rel_err(xg.grad, xo.grad) < TOL
    for (n, p), (_, po) in zip(m.named_parameters(), om.named_parameters()):
        assert p.grad is not None, n
        assert rel_err(p.grad, po.grad) < 2e-3, n


@pytest.mark.gpu
@pytest.mark.parametrize("C,Co,wg,og,k,stride,pad,dil,use_mask,use_bias", [
    (16, 16, 16, 1, (5, 5), 1, 2, 1, False, False),       # depthwise, as conv0 of the 2D block (register-accumulated weight gradient)
    (16, 16, 16, 1, (7, 7), 1, 9, 3, False, False),       # depthwise dilated, as conv_spatial
    (40, 40, 40, 1, (3, 3), 1, 1, 1, False, True),        # depthwise 3x3, channels not a multiple of 32
    (8, 8, 1, 1, (3, 3), 1, 1, 1, False, True),           # dense
    (16, 8, 2, 2, (3, 5), (1, 2), (2, 3), (2, 1), True, True),   # weight groups + offset groups + mask + stride / dilation
    (64, 64, 64, 2, (3, 3), 2, 1, 1, True, False),        # depthwise + 2 offset groups of 32 channels (warp-uniform reduction) + mask
])
def test_deform_conv2d_backward_vs_torchvision_autograd(C, Co, wg, og, k, stride, pad, dil, use_mask, use_bias):
    import torchvision
    import deformablelka_b200 as dl
    from oracle import oracle as o
    torch.manual_seed(5)
    B, H, W = 2, 13, 11
    kh, kw = k
    sh, sw = o._pair(stride); ph, pw = o._pair(pad); dh, dw = o._pair(dil)
    Ho, Wo = o.out_extent(H, ph, dh, kh, sh), o.out_extent(W, pw, dw, kw, sw)
    x = torch.randn(B, C, H, W, requires_grad=True)
    w = (torch.randn(Co, C // wg, kh, kw) * 0.3).requires_grad_()
    b = torch.randn(Co, requires_grad=True) if use_bias else None
    off = (torch.randn(B, og * 2 * kh * kw, Ho, Wo) * 2).requires_grad_()
    mask = torch.rand(B, og * kh * kw, Ho, Wo, requires_grad=True) if use_mask else None
    gout = torch.randn(B, Co, Ho, Wo)
    torchvision.ops.deform_conv2d(x, off, w, b, stride, pad, dil, mask).backward(gout)
    dev = "cuda:0"
    xg, wg_, og_ = (t.detach().to(dev).requires_grad_() for t in (x, w, off))
    bg = None if b is None else b.detach().to(dev).requires_grad_()
    mg = None if mask is None else mask.detach().to(dev).requires_grad_()
    from deformablelka_b200.deformable_LKA import deform_conv2d_autograd
    y = deform_conv2d_autograd(xg, og_, wg_, bg, stride, pad, dil, mg)
    y.backward(gout.to(dev))
    pairs = [("grad_input", xg.grad, x.grad), ("grad_weight", wg_.grad, w.grad), ("grad_offset", og_.grad, off.grad)]
    if use_bias:
        pairs.append(("grad_bias", bg.grad, b.grad))
    if use_mask:
        pairs.append(("grad_mask", mg.grad, mask.grad))
    for name, got, want in pairs:
        assert got.shape == want.shape, name
        assert rel_err(got, want) < TOL, name


@pytest.mark.gpu
def test_block2d_backward_on_gpu_vs_oracle(oracle):
    import deformablelka_b200 as dl
    torch.manual_seed(6)
    C, H, W = 32, 12, 10
    m = dl.deformable_LKA_Attention(C)
    om = oracle.deformable_LKA_Attention(C)
    om.load_state_dict(m.state_dict())
    x = torch.randn(2, C, H, W)
    xo = x.clone().requires_grad_()
    om(xo).square().sum().backward()
    m = m.to("cuda:0")
    xg = x.to("cuda:0").requires_grad_()
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        y = m(xg)
        assert y.grad_fn is not None
        y.square().sum().backward()
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    with torch.no_grad():
        assert rel_err(m(xg.detach()), om(x)) < TOL      # the fused inference call computes the same function
    assert rel_err(xg.grad, xo.grad) < TOL
    for (n, p), (_, po) in zip(m.named_parameters(), om.named_parameters()):
        assert p.grad is not None, n
        assert rel_err(p.grad, po.grad) < 2e-3, n


def test_graphed_call_refuses_cpu_tensors():
    """deformablelka_b200.graphs.GraphedCall captures a CUDA graph: CPU arguments fail loudly, like every other entry."""
    import deformablelka_b200 as dl
    with pytest.raises(RuntimeError, match="CPU"):
        dl.GraphedCall(lambda t: t, torch.zeros(2))
    with pytest.raises(RuntimeError, match="CPU"):
        dl.GraphedCall(lambda: None)

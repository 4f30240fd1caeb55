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


def test_fused_2d_entries_refuse_when_grad_is_required():
    import deformablelka_b200 as dl
    m = dl.deformable_LKA_Attention(8)
    blk = dl.deformableLKABlock(8)
    x = torch.randn(1, 8, 5, 5)
    with mock.patch.object(torch.Tensor, "is_cuda", new_callable=mock.PropertyMock, return_value=True):
        for call in (lambda: m(x), lambda: m.spatial_gating_unit(x), lambda: m.spatial_gating_unit.conv0(x),
                     lambda: blk(torch.randn(1, 25, 8), 5, 5)):
            with pytest.raises(RuntimeError, match="gradient is required"):
                call()
        m.requires_grad_(False)
        with pytest.raises(RuntimeError, match="gradient is required"):   # the input still asks for one
            m(x.clone().requires_grad_())
    # CPU tensors are refused either way (no CPU path)
    with pytest.raises(RuntimeError, match="CPU"):
        dl.deformable_LKA_Attention(8)(x)
    with torch.no_grad(), pytest.raises(RuntimeError, match="CPU"):
        dl.deformable_LKA_Attention(8)(x)
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
def test_fused_2d_refuses_on_gpu():
    import deformablelka_b200 as dl
    m = dl.deformable_LKA_Attention(8).to("cuda:0")
    x = torch.randn(1, 8, 6, 6, device="cuda:0")
    with pytest.raises(RuntimeError, match="gradient is required"):
        m(x)
    with torch.no_grad():
        assert m(x).shape == x.shape

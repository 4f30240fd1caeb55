"""CUDA-graph replay of library calls (deformablelka_b200.graphs.GraphedCall): a captured call must reproduce the eager call
bit for bit (same kernels, same order), follow new inputs, and refuse mismatched arguments."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def dl():
    import deformablelka_b200 as dl
    return dl


def _block3d(dl, C):
    torch.manual_seed(7)
    m = dl.LKA_Attention3d_deform(C)
    with torch.no_grad():
        m.spatial_gating_unit.deform_conv.conv_offset.weight.normal_(0, 0.05)
        m.spatial_gating_unit.deform_conv.conv_offset.bias.uniform_(-1, 1)
    return m.to(DEV).eval()


@pytest.mark.parametrize("C,S", [(256, 4), (64, 8), (32, 12)])
def test_graphed_block3d_equals_eager_and_follows_inputs(dl, C, S):
    m = _block3d(dl, C)
    x1 = torch.randn(2, S * S * S, C, device=DEV)
    x2 = torch.randn(2, S * S * S, C, device=DEV)
    with torch.no_grad():
        e1, e2 = m(x1, 2, C, S, S, S).clone(), m(x2, 2, C, S, S, S).clone()
    g = dl.GraphedCall(m, x1, 2, C, S, S, S)
    assert torch.equal(g(x1, 2, C, S, S, S), e1)
    assert torch.equal(g(x2, 2, C, S, S, S), e2)      # static input buffer refreshed before the replay
    assert torch.equal(g(x1, 2, C, S, S, S), e1)
    with pytest.raises(RuntimeError, match="does not match"):
        g(torch.randn(1, S * S * S, C, device=DEV), 2, C, S, S, S)
    with pytest.raises(RuntimeError, match="differs"):
        g(x1, 2, C, S, S, S + 1)


def test_graphed_block2d_equals_eager(dl):
    torch.manual_seed(8)
    m = dl.deformable_LKA_Attention(32).to(DEV).eval()
    x1, x2 = torch.randn(3, 32, 14, 14, device=DEV), torch.randn(3, 32, 14, 14, device=DEV)
    with torch.no_grad():
        e1, e2 = m(x1).clone(), m(x2).clone()
    g = dl.GraphedCall(m, x1)
    assert torch.equal(g(x2), e2)
    assert torch.equal(g(x1), e1)


def test_recapture_sees_new_parameters(dl):
    m = _block3d(dl, 32)
    x = torch.randn(1, 216, 32, device=DEV)
    g = dl.GraphedCall(m, x, 1, 32, 6, 6, 6)
    with torch.no_grad():
        m.proj_2.weight.mul_(0.5)
        want = m(x, 1, 32, 6, 6, 6).clone()
    g.recapture()
    assert torch.equal(g(x, 1, 32, 6, 6, 6), want)

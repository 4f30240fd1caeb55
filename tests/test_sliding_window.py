"""Rest of row N4: the sliding-window predictor (host orchestration, runs on CPU) against the numpy restatement of the
reference's SegmentationNetwork tiled prediction, and its helpers against the reference's own printed examples."""
import numpy as np
import pytest
import torch

from deformablelka_b200 import sliding_window as sw


def test_steps_match_reference_examples():
    # the reference's own examples (neural_network.py:815-820, and the worked example in the comment at :271-272)
    assert sw.compute_steps_for_sliding_window((64,), (110,), 0.5) == [[0, 23, 46]]
    assert sw.compute_steps_for_sliding_window((30, 224, 224), (60, 448, 224), 1) == [[0, 30], [0, 224], [0]]
    st = sw.compute_steps_for_sliding_window((30, 224, 224), (162, 529, 529), 0.5)
    assert st[0][0] == 0 and st[0][-1] == 162 - 30 and st[1][-1] == 529 - 224
    assert all(b - a <= 15 for a, b in zip(st[0], st[0][1:]))          # never further apart than patch * step_size
    with pytest.raises(AssertionError):
        sw.compute_steps_for_sliding_window((8, 8, 8), (16, 16, 16), 0.0)


def test_gaussian_map_properties():
    g = sw.gaussian_importance_map((8, 12, 10))
    assert g.dtype == np.float32 and g.shape == (8, 12, 10)
    assert g.max() == 1.0 and g[4, 6, 5] == 1.0 and g.min() > 0


class _Net(torch.nn.Module):
    """A small asymmetric network: flips must be undone exactly for the mirrored average to be right."""

    def __init__(self, cin, classes):
        super().__init__()
        torch.manual_seed(3)
        self.c1 = torch.nn.Conv3d(cin, 6, 3, padding=1)
        self.c2 = torch.nn.Conv3d(6, classes, (3, 1, 3), padding=(1, 0, 1))

    def forward(self, x):
        return self.c2(torch.tanh(self.c1(x)))


@pytest.mark.parametrize("shape,patch,step,mirror,axes,gauss", [
    ((2, 20, 17, 13), (8, 8, 8), 0.5, True, (0, 1, 2), True),
    ((1, 9, 30, 11), (8, 12, 8), 0.75, True, (1, 2), True),
    ((2, 16, 16, 16), (8, 8, 8), 1.0, False, (0, 1, 2), False),
    ((1, 5, 6, 20), (8, 8, 8), 0.5, True, (0,), True),          # volume smaller than the patch on two axes: padded, then cropped
    ((1, 8, 8, 8), (8, 8, 8), 0.5, True, (0, 1, 2), True),      # a single tile: no Gaussian weighting (reference :321)
])
def test_predict_3d_tiled_vs_oracle(oracle, shape, patch, step, mirror, axes, gauss):
    torch.manual_seed(5)
    classes = 4
    net = _Net(shape[0], classes).eval()
    x = torch.randn(*shape)
    seg, probs = sw.predict_3d_tiled(net, x, patch, classes, step, mirror, axes, gauss)
    seg_o, probs_o = oracle.sliding_window_predict_oracle(net, x.numpy(), patch, classes, step, mirror, axes, gauss)
    assert probs.shape == (classes,) + tuple(shape[1:]) and seg.shape == tuple(shape[1:])
    np.testing.assert_allclose(probs, probs_o, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(probs.sum(0), 1.0, atol=1e-4)         # weighted average of softmax outputs
    # argmax may differ only where two classes are numerically tied
    diff = seg != seg_o
    if diff.any():
        top2 = np.sort(probs_o, 0)[-2:]
        assert np.all((top2[1] - top2[0])[diff] < 1e-5)


def test_regions_class_order():
    net = _Net(1, 3).eval()
    x = torch.randn(1, 10, 9, 8)
    seg, probs = sw.predict_3d_tiled(net, x, (8, 8, 8), 3, regions_class_order=(1, 2, 3))
    exp = np.zeros(probs.shape[1:], dtype=np.float32)
    for i, c in enumerate((1, 2, 3)):
        exp[probs[i] > 0.5] = c
    assert np.array_equal(seg, exp)

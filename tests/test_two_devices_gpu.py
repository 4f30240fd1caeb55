"""One process driving TWO GPUs (what nn.DataParallel, the reference 2D trainers' only multi-GPU mode, does): per-device state of
the library -- the shared-memory opt-in of every kernel, the device check, the helper streams of the host path -- must be keyed
by device.  Skipped on single-GPU boxes."""
import threading

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dl():
    import deformablelka_b200 as dl
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible GPUs")
    return dl


def _block(dl, C, dev):
    torch.manual_seed(11)
    m = dl.LKA_Attention3d_deform(C)
    with torch.no_grad():
        m.spatial_gating_unit.deform_conv.conv_offset.weight.normal_(0, 0.05)
        m.spatial_gating_unit.deform_conv.conv_offset.bias.uniform_(-1, 1)
    return m.to(dev).eval()


def test_second_device_gets_its_own_smem_opt_in_and_matches_the_first(dl):
    x = torch.randn(1, 6 * 7 * 9, 32)
    outs = []
    for dev in ("cuda:0", "cuda:1", "cuda:0"):          # device 1 is used AFTER device 0 configured every kernel
        m = _block(dl, 32, dev)
        with torch.no_grad():
            outs.append(m(x.to(dev), 1, 32, 6, 7, 9).cpu())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


def test_two_devices_from_two_host_threads(dl):
    x = torch.randn(2, 8 * 8 * 8, 64)
    res, err = {}, []

    # the modules are built one after the other on this thread (torch's CPU generator is process-wide: two threads seeding and
    # drawing from it concurrently would initialise two DIFFERENT modules)
    mods = {d: _block(dl, 64, d) for d in ("cuda:0", "cuda:1")}

    def work(dev):
        try:
            m = mods[dev]
            with torch.no_grad():
                for _ in range(3):
                    y = m(x.to(dev), 2, 64, 8, 8, 8)
                res[dev] = y.cpu()
        except Exception as e:   # noqa: BLE001
            err.append((dev, repr(e)))

    th = [threading.Thread(target=work, args=(d,)) for d in ("cuda:0", "cuda:1")]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not err, err
    assert torch.equal(res["cuda:0"], res["cuda:1"])


def test_host_path_on_the_second_device(dl):
    m = _block(dl, 32, "cuda:1")
    xh = torch.randn(2, 5 * 6 * 7, 32).pin_memory()
    with torch.no_grad():
        want = m(xh.to("cuda:1"), 2, 32, 5, 6, 7).cpu()
        got = m.forward_host(xh, 2, 32, 5, 6, 7)
    assert torch.equal(torch.as_tensor(got).cpu().reshape(want.shape), want)


def test_2d_block_and_operator_on_the_second_device(dl):
    torch.manual_seed(12)
    m0 = dl.deformable_LKA_Attention(32).eval()
    x = torch.randn(2, 32, 14, 14)
    with torch.no_grad():
        y0 = m0.to("cuda:0")(x.to("cuda:0")).cpu()
        y1 = m0.to("cuda:1")(x.to("cuda:1")).cpu()
    assert torch.equal(y0, y1)

"""Golden vectors made by the UNMODIFIED reference sources (tests/golden/make_golden_blocks.py, ref_import.py) for rows
a8 / a9 and N1 / N3 / N4 of SURVEY.md section 8.

* CPU tests (``-m "not gpu"``): the oracle's block restatements against the goldens -- this is what pins the oracle that the
  larger GPU parity tests use as their reference;
* GPU tests: the product modules (one library call each) against the same goldens.

Tolerance: 1e-3 of the output range (north_star) for the GPU path; the oracle runs the same stock layers in the same order
as the reference, so it is held to 1e-5.
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

TOL = 1e-3
ORACLE_TOL = 1e-5
DEV = "cuda:0"


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    sd, ins, outs, meta = {}, {}, {}, {}
    for k in z.files:
        kind, key = k.split(".", 1)
        v = z[k]
        if kind == "sd":
            t = torch.from_numpy(v)
            sd[key] = t.float() if t.is_floating_point() else t
        elif kind == "in":
            ins[key] = torch.from_numpy(v)
        elif kind == "out":
            outs[key] = torch.from_numpy(v)
        else:
            meta[key] = v.item() if v.ndim == 0 else v
    return sd, ins, outs, meta


def rel_err(got, ref):
    got = got.detach().float().cpu(); ref = ref.detach().float().cpu()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


LKABLOCKS = ["ref2d_lkablock_c16", "ref2d_lkablock_c32"]
EXPANDS = ["ref2d_patchexpand_d32", "ref2d_patchexpand_d96", "ref2d_finalexpand_d32"]
DECODERS = ["ref2d_decoder_d16_mid", "ref2d_decoder_d16_last"]
TBLOCKS = ["ref3d_synapse_tblock_c32", "ref3d_synapse_tblock_c8", "ref3d_acdc_tblock_c32"]


def _dim_of(name):
    return int(name.rsplit("_", 1)[1][1:]) if not name.startswith("ref2d_decoder") else int(name.split("_")[2][1:])


# ============================================================================== CPU: the oracle against the reference
@pytest.mark.parametrize("name", LKABLOCKS)
def test_oracle_lkablock_vs_reference(oracle, name):
    sd, ins, outs, meta = load(name)
    m = oracle.deformableLKABlock(_dim_of(name)).eval()
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        assert rel_err(m(ins["x"], meta["H"], meta["W"]), outs["y"]) < ORACLE_TOL


@pytest.mark.parametrize("name", EXPANDS)
def test_oracle_patch_expand_vs_reference(oracle, name):
    sd, ins, outs, meta = load(name)
    cls = oracle.FinalPatchExpand_X4 if "final" in name else oracle.PatchExpand
    m = cls((meta["H"], meta["W"]), _dim_of(name)).eval()
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        assert rel_err(m(ins["x"]), outs["y"]) < ORACLE_TOL


@pytest.mark.parametrize("name", DECODERS)
def test_oracle_decoder_layer_vs_reference(oracle, name):
    sd, ins, outs, meta = load(name)
    dim = _dim_of(name)
    m = oracle.MyDecoderLayer((meta["H"], meta["W"]), [dim] * 5, 1, "mix_skip", n_class=9, is_last=bool(meta["is_last"])).eval()
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        assert rel_err(m(ins["x1"], ins["x2"]), outs["y"]) < ORACLE_TOL
        assert rel_err(m(ins["x1"]), outs["y_noskip"]) < ORACLE_TOL


def _oracle_3d(oracle, name, sd, C):
    acdc = "acdc" in name
    attn = (oracle.LKA_Attention3d_deform_ACDC if acdc else oracle.LKA_Attention3d_deform)(C).eval()
    attn.load_state_dict({k[len("epa_block."):]: v for k, v in sd.items() if k.startswith("epa_block.")}, strict=True)
    return attn


@pytest.mark.parametrize("name", TBLOCKS)
def test_oracle_block3d_vs_reference(oracle, name):
    """LKA3d_deform, LKA_Attention3d_deform and the whole TransformerBlock_3D_single_deform_LKA as composed by the reference's
    own transformerblock.py / dynunet_block.py."""
    sd, ins, outs, meta = load(name)
    C = _dim_of(name)
    H, W, D = meta["H"], meta["W"], meta["D"]
    attn = _oracle_3d(oracle, name, sd, C)
    norm = torch.nn.LayerNorm(C); norm.load_state_dict({"weight": sd["norm.weight"], "bias": sd["norm.bias"]})
    res = oracle.UnetResBlock3D(C).eval()
    res.load_state_dict({k[len("conv51."):]: v for k, v in sd.items() if k.startswith("conv51.")}, strict=True)
    conv8 = torch.nn.Conv3d(C, C, 1); conv8.load_state_dict({"weight": sd["conv8.1.weight"], "bias": sd["conv8.1.bias"]})
    with torch.no_grad():
        assert rel_err(attn.spatial_gating_unit(ins["xv"]), outs["y_lka"]) < ORACLE_TOL
        assert rel_err(attn(ins["tokens"], 2, C, H, W, D), outs["y_attn"]) < ORACLE_TOL
        got = oracle.transformer3d_block(norm, sd["gamma"], attn, sd["pos_embed"], res, conv8, ins["x"])
        assert rel_err(got, outs["y"]) < ORACLE_TOL


def test_oracle_acdc_attention_c128_vs_reference(oracle):
    sd, ins, outs, meta = load("ref3d_acdc_attn_c128")
    m = oracle.LKA_Attention3d_deform_ACDC(128).eval()
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        assert rel_err(m(ins["tokens"], 1, 128, meta["H"], meta["W"], meta["D"]), outs["y_attn"]) < ORACLE_TOL


def test_oracle_c1_block2d_vs_reference(oracle):
    """BASELINE.json configs[0]: the reference's own CPU case, 1x64x224x224, against the oracle on the full tensor (the golden
    holds the reference output on a stride-(5,3) lattice plus per-channel sums of the whole tensor)."""
    z = np.load(os.path.join(GOLDEN, "ref2d_c1_attn_c64.npz"))
    sd = {k[3:]: torch.from_numpy(z[k]).float() for k in z.files if k.startswith("sd.")}
    x = torch.randn(1, 64, 224, 224, generator=torch.Generator().manual_seed(int(z["meta.x_seed"])))
    assert torch.equal(x.flatten()[:16], torch.from_numpy(z["meta.x_head"]))
    m = oracle.deformable_LKA_Attention(64).eval()
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        y = m(x)
    ref_sub = torch.from_numpy(z["out.y_sub"])
    assert rel_err(y[:, :, ::5, ::3], ref_sub) < ORACLE_TOL
    chan = y.double().sum((0, 2, 3))
    assert ((chan - torch.from_numpy(z["out.y_chan_sum"])).abs().max() / (224 * 224 * ref_sub.abs().max())).item() < 1e-6


def test_sliding_window_helpers_vs_reference():
    """compute_steps_for_sliding_window / gaussian_importance_map (host logic of the product) against the reference's own
    static helpers (neural_network.py:250-290), evaluated in the build container."""
    from deformablelka_b200 import sliding_window as sw
    z = np.load(os.path.join(GOLDEN, "ref3d_sliding_window_helpers.npz"))
    i = 0
    while f"steps{i}.patch" in z.files:
        patch, image, step = tuple(z[f"steps{i}.patch"]), tuple(z[f"steps{i}.image"]), float(z[f"steps{i}.step"])
        got = sw.compute_steps_for_sliding_window(patch, image, step)
        for ax in range(3):
            assert got[ax] == list(z[f"steps{i}.ax{ax}"]), (patch, image, step)
        i += 1
    assert i >= 5
    i = 0
    while f"gauss{i}.patch" in z.files:
        g = sw.gaussian_importance_map(tuple(z[f"gauss{i}.patch"]))
        ref = z[f"gauss{i}.map"]
        assert g.dtype == ref.dtype and np.array_equal(g, ref)
        i += 1
    assert i >= 3


# ============================================================================== GPU: the product against the reference
@pytest.fixture(params=["fp32", "bf16x3"])
def math(request, monkeypatch):
    monkeypatch.setenv("DLKA_MATH", request.param)
    return request.param


@pytest.mark.gpu
@pytest.mark.parametrize("name", LKABLOCKS)
def test_lkablock_vs_reference(name, math):
    import deformablelka_b200 as dl
    sd, ins, outs, meta = load(name)
    m = dl.deformableLKABlock(_dim_of(name))
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        got = m.to(DEV)(ins["x"].to(DEV), meta["H"], meta["W"])
    assert rel_err(got, outs["y"]) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("name", EXPANDS)
def test_patch_expand_vs_reference(name, math):
    import deformablelka_b200 as dl
    sd, ins, outs, meta = load(name)
    cls = dl.FinalPatchExpand_X4 if "final" in name else dl.PatchExpand
    m = cls((meta["H"], meta["W"]), _dim_of(name))
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        got = m.to(DEV)(ins["x"].to(DEV))
    assert rel_err(got, outs["y"]) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("name", DECODERS)
def test_decoder_layer_vs_reference(name, math):
    import deformablelka_b200 as dl
    sd, ins, outs, meta = load(name)
    dim = _dim_of(name)
    m = dl.MyDecoderLayer((meta["H"], meta["W"]), [dim] * 5, 1, "mix_skip", n_class=9, is_last=bool(meta["is_last"]))
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV)
    with torch.no_grad():
        got = m(ins["x1"].to(DEV), ins["x2"].to(DEV))
        got0 = m(ins["x1"].to(DEV))
    assert rel_err(got, outs["y"]) < TOL
    assert rel_err(got0, outs["y_noskip"]) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("name", TBLOCKS)
def test_block3d_vs_reference(name, math):
    import deformablelka_b200 as dl
    from deformablelka_b200 import acdc
    sd, ins, outs, meta = load(name)
    C = _dim_of(name)
    H, W, D = meta["H"], meta["W"], meta["D"]
    cls = acdc.TransformerBlock_3D_single_deform_LKA if "acdc" in name else dl.TransformerBlock_3D_single_deform_LKA
    m = cls(H * W * D, C, C, 4, pos_embed=True)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval()
    with torch.no_grad():
        got = m(ins["x"].to(DEV))
        got_attn = m.epa_block(ins["tokens"].to(DEV), 2, C, H, W, D)
        got_lka = m.epa_block.spatial_gating_unit(ins["xv"].to(DEV))
    assert rel_err(got, outs["y"]) < TOL
    assert rel_err(got_attn, outs["y_attn"]) < TOL
    assert rel_err(got_lka, outs["y_lka"]) < TOL


@pytest.mark.gpu
def test_acdc_attention_c128_vs_reference(math):
    from deformablelka_b200 import acdc
    sd, ins, outs, meta = load("ref3d_acdc_attn_c128")
    m = acdc.LKA_Attention3d_deform(128)
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        got = m.to(DEV)(ins["tokens"].to(DEV), 1, 128, meta["H"], meta["W"], meta["D"])
    assert rel_err(got, outs["y_attn"]) < TOL

"""Golden vectors for the enclosing blocks and the 3D block (rows a8 / a9), produced by
running the UNMODIFIED reference sources on CPU in the build container:

    python tests/golden/make_golden_blocks.py

2D  2D/networks/MaxViT_deform_LKA.py:142-189,488-620   deformableLKABlock, PatchExpand, FinalPatchExpand_X4, MyDecoderLayer
3D  3D/d_lka_former/network_architecture/{synapse,acdc}/transformerblock.py   LKA3d_deform, LKA_Attention3d_deform,
    TransformerBlock_3D_single_deform_LKA (with dynunet_block.UnetResBlock from the reference tree)
    3D/d_lka_former/network_architecture/neural_network.py:250-290   sliding-window steps and Gaussian importance map

How the missing third-party imports are served is documented in tests/golden/ref_import.py (the 3D operator ``D3D`` is the
C oracle restatement, which the GPU tests pin against the reference's compiled extension).  Every fixture stores the
state_dict, the inputs and the outputs; parameters the reference initialises to (near) zero -- layer scales 1e-2, gamma 1e-6,
zero conv_offset, fresh BatchNorm statistics -- are randomised first so that every branch contributes to the output.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_import  # noqa: E402


def fp16_exact_(module):
    """Round every floating-point parameter / buffer to an fp16-representable value BEFORE the reference runs, so the
    fixture can store the state_dict in half the bytes and still reproduce the reference's fp32 inputs bit for bit."""
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            if t.is_floating_point():
                t.copy_(t.half().float())


def save(name, module, inputs: dict, outputs: dict, meta: dict = None):
    out = {}
    for k, v in module.state_dict().items():
        v = v.detach()
        if v.is_floating_point():
            assert torch.equal(v.half().float(), v), k   # fp16_exact_ was applied
            v = v.half()
        out["sd." + k] = v.numpy()
    for k, v in inputs.items():
        out["in." + k] = v.detach().numpy()
    for k, v in outputs.items():
        out["out." + k] = v.detach().numpy()
    for k, v in (meta or {}).items():
        out["meta." + k] = np.asarray(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB", {k: tuple(v.shape) for k, v in outputs.items()})


def liven_2d_block(blk, scale, g):
    with torch.no_grad():
        blk.layer_scale_1.uniform_(0.2, 1.0, generator=g); blk.layer_scale_2.uniform_(0.2, 1.0, generator=g)
        blk.norm1.weight.uniform_(0.5, 1.5, generator=g); blk.norm1.bias.normal_(0, 0.1, generator=g)
        blk.norm2.weight.uniform_(0.5, 1.5, generator=g); blk.norm2.bias.normal_(0, 0.1, generator=g)
        for n, mod in blk.named_modules():
            if n.endswith("offset_net"):
                mod.weight.mul_(scale); mod.bias.mul_(scale)


def randomize_offsets(module, g, std=0.05, bias_range=1.0):
    with torch.no_grad():
        for m in module.modules():
            if hasattr(m, "conv_offset"):
                m.conv_offset.weight.normal_(0, std, generator=g)
                m.conv_offset.bias.uniform_(-bias_range, bias_range, generator=g)


def liven_3d_block(m, g):
    randomize_offsets(m, g)
    with torch.no_grad():
        m.gamma.uniform_(0.2, 1.0, generator=g)
        if m.pos_embed is not None:
            m.pos_embed.normal_(0, 0.5, generator=g)
        m.norm.weight.uniform_(0.5, 1.5, generator=g); m.norm.bias.normal_(0, 0.1, generator=g)
        for bn in (m.conv51.norm1, m.conv51.norm2):
            bn.running_mean.normal_(0, 0.3, generator=g); bn.running_var.uniform_(0.5, 2.0, generator=g)
            bn.weight.uniform_(0.5, 1.5, generator=g); bn.bias.normal_(0, 0.2, generator=g)


def main():
    torch.set_num_threads(8)
    # ------------------------------------------------------------------ 2D decoder-side blocks
    ref = ref_import.import_reference_maxvit()
    for dim, hw, scale in ((16, (6, 5), 1.0), (32, (9, 7), 4.0)):
        torch.manual_seed(100 + dim)
        g = torch.Generator().manual_seed(dim)
        m = ref.deformableLKABlock(dim).eval()
        liven_2d_block(m, scale, g)
        fp16_exact_(m)
        x = torch.randn(2, hw[0] * hw[1], dim)
        with torch.no_grad():
            y = m(x, *hw)
        save(f"ref2d_lkablock_c{dim}", m, {"x": x}, {"y": y}, {"H": hw[0], "W": hw[1]})
    for dim, hw, cls, tag in ((32, (7, 5), ref.PatchExpand, "patchexpand"), (96, (4, 3), ref.PatchExpand, "patchexpand"),
                              (32, (5, 4), ref.FinalPatchExpand_X4, "finalexpand")):
        torch.manual_seed(200 + dim)
        g = torch.Generator().manual_seed(dim)
        m = cls(hw, dim).eval()
        with torch.no_grad():
            m.norm.weight.uniform_(0.5, 1.5, generator=g); m.norm.bias.normal_(0, 0.2, generator=g)
        fp16_exact_(m)
        x = torch.randn(2, hw[0] * hw[1], dim)
        with torch.no_grad():
            y = m(x)
        save(f"ref2d_{tag}_d{dim}", m, {"x": x}, {"y": y}, {"H": hw[0], "W": hw[1]})
    for dim, hw, last in ((16, (6, 5), False), (16, (4, 6), True)):
        torch.manual_seed(300 + dim + int(last))
        g = torch.Generator().manual_seed(dim + int(last))
        m = ref.MyDecoderLayer(hw, [dim] * 5, 1, "mix_skip", n_class=9, is_last=last).eval()
        for blk in (m.layer_lka_1, m.layer_lka_2):
            liven_2d_block(blk, 1.0, g)
        with torch.no_grad():
            m.x1_linear.bias.normal_(0, 0.2, generator=g)
            if last:
                m.last_layer.bias.normal_(0, 0.2, generator=g)
        fp16_exact_(m)
        x1, x2 = torch.randn(2, hw[0] * hw[1], dim), torch.randn(2, hw[0], hw[1], dim)
        with torch.no_grad():
            y = m(x1, x2)
            y0 = m(x1)
        save(f"ref2d_decoder_d{dim}_{'last' if last else 'mid'}", m, {"x1": x1, "x2": x2}, {"y": y, "y_noskip": y0},
             {"H": hw[0], "W": hw[1], "is_last": int(last)})

    # ------------------------------------------------------------------ 3D blocks (synapse + ACDC)
    for net, dims in (("synapse", ((32, (6, 5, 8)), (8, (5, 4, 6)))), ("acdc", ((32, (4, 9, 8)),))):
        tb = ref_import.import_reference_3d(net)
        for C, (H, W, D) in dims:
            torch.manual_seed(400 + C)
            g = torch.Generator().manual_seed(C)
            N = H * W * D
            m = tb.TransformerBlock_3D_single_deform_LKA(N, C, C, 4, pos_embed=True).eval()
            liven_3d_block(m, g)
            fp16_exact_(m)
            x = torch.randn(2, C, H, W, D)
            tok = torch.randn(2, N, C)
            xv = torch.randn(2, C, H, W, D)
            with torch.no_grad():
                y = m(x)
                y_attn = m.epa_block(tok, 2, C, H, W, D)
                y_lka = m.epa_block.spatial_gating_unit(xv)
            save(f"ref3d_{net}_tblock_c{C}", m, {"x": x, "tokens": tok, "xv": xv},
                 {"y": y.contiguous(), "y_attn": y_attn.contiguous(), "y_lka": y_lka}, {"H": H, "W": W, "D": D})

    # ACDC dim 128 selects the (3,5,5)-dil-(1,3,3) stencil (acdc/transformerblock.py:224-229): attention block only (the
    # transformer tail's two 128x128x27 convs would triple the fixture)
    tb = ref_import.import_reference_3d("acdc")
    torch.manual_seed(528)
    g = torch.Generator().manual_seed(128)
    C, (H, W, D) = 128, (3, 6, 5)
    m = tb.LKA_Attention3d_deform(C).eval()
    randomize_offsets(m, g)
    fp16_exact_(m)
    tok = torch.randn(1, H * W * D, C)
    with torch.no_grad():
        y_attn = m(tok, 1, C, H, W, D)
    save("ref3d_acdc_attn_c128", m, {"tokens": tok}, {"y_attn": y_attn.contiguous()}, {"H": H, "W": W, "D": D})

    # ------------------------------------------------------------------ BASELINE.json configs[0] (C1): the reference's own
    # CPU-runnable case, deformable_LKA_Attention / deformable_LKA on ONE 1x64x224x224 tensor.  The 12.8 MB input is
    # regenerated from the seed by the test (checksums stored); the output is stored on a stride-(5,3) lattice (+ per-channel
    # float64 sums of the full tensor).
    ref2d = ref_import.import_reference_2d_module()
    torch.manual_seed(1234)
    m = ref2d.deformable_LKA_Attention(64).eval()
    fp16_exact_(m)
    gx = torch.Generator().manual_seed(4321)
    x = torch.randn(1, 64, 224, 224, generator=gx)
    with torch.no_grad():
        y = m(x)
        y_lka = m.spatial_gating_unit(x)
    save("ref2d_c1_attn_c64", m, {}, {"y_sub": y[:, :, ::5, ::3].contiguous(), "y_lka_sub": y_lka[:, :, ::5, ::3].contiguous(),
                                      "y_chan_sum": y.double().sum((0, 2, 3)), "y_lka_chan_sum": y_lka.double().sum((0, 2, 3))},
         {"x_seed": 4321, "x_sum": float(x.double().sum()), "x_abs_sum": float(x.double().abs().sum()),
          "x_head": x.flatten()[:16].numpy()})

    # ------------------------------------------------------------------ sliding-window helpers
    sn = ref_import.import_reference_segmentation_network().SegmentationNetwork
    out = {}
    cases = [((64, 128, 128), (100, 200, 300), 0.5), ((8, 12, 10), (8, 12, 10), 0.5), ((8, 8, 8), (21, 9, 30), 0.25),
             ((16, 16, 16), (17, 40, 16), 1.0), ((4, 6, 5), (11, 13, 7), 0.5)]
    for i, (patch, image, step) in enumerate(cases):
        steps = sn._compute_steps_for_sliding_window(patch, image, step)
        out[f"steps{i}.patch"] = np.asarray(patch); out[f"steps{i}.image"] = np.asarray(image); out[f"steps{i}.step"] = np.asarray(step)
        for ax in range(3):
            out[f"steps{i}.ax{ax}"] = np.asarray(steps[ax], dtype=np.int64)
    for i, patch in enumerate(((8, 12, 10), (4, 6, 5), (16, 16, 16))):
        out[f"gauss{i}.patch"] = np.asarray(patch)
        out[f"gauss{i}.map"] = sn._get_gaussian(patch, sigma_scale=1. / 8)
    path = os.path.join(HERE, "ref3d_sliding_window_helpers.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()

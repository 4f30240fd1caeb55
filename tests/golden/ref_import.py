"""Import the UNMODIFIED reference sources in the build container (``/root/reference`` does not exist on the GPU box).

The reference files import packages that are not installed here (fvcore, timm, monai, batchgenerators, the compiled
``D3D`` extension).  Each helper injects the smallest possible stand-in into ``sys.modules`` and then imports the
reference file itself, so everything the golden vectors pin -- module composition, reshapes / permutes, parameter
names, residual / layer-scale arithmetic -- is executed from the reference's own source:

  fvcore.nn.FlopCountAnalysis      unused at run time (2D/deformable_LKA/deformable_LKA.py:160)
  timm.models.layers.DropPath      identity for drop_path = 0 (the only value the decoder constructs)
  networks.merit_lib.networks      the MaxViT encoder (needs timm + a checkpoint download): out of scope, stubbed
  monai 0.8.1 leaf layers          ``Convolution(conv_only=True)`` = a bias-free nn.Conv3d under ``.conv`` with the padding
                                   the reference computes (dynunet_block.py:217-248), ``get_norm_layer("batch")`` =
                                   nn.BatchNorm3d, ``get_act_layer(("leakyrelu", kw))`` = nn.LeakyReLU(**kw): monai's published
                                   behaviour for exactly the arguments UnetResBlock passes (dynunet_block.py:42-56)
  D3D                              the CUDA-only extension (3D/dcn/src/deform_conv.h:46).  ``deform_conv_forward`` is
                                   served by the C oracle restatement (oracle/dlka_oracle.c), which is itself pinned against
                                   the reference's compiled D3D on the GPU box (tests/test_ref_d3d_gpu.py).  The block-level
                                   goldens therefore pin the COMPOSITION from reference source; the operator arithmetic
                                   is pinned separately.
"""
import os
import sys
import types

import torch.nn as nn

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _stub(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        sys.modules[name] = m
    m.__dict__.update(attrs)
    return m


def available() -> bool:
    return os.path.isdir(REF)


def import_reference_2d_module():
    """2D/deformable_LKA/deformable_LKA.py (deformable_LKA, deformable_LKA_Attention)."""
    _stub("fvcore"); _stub("fvcore.nn", FlopCountAnalysis=object)
    import importlib.util
    name = "_reference_2d_deformable_LKA"   # loaded by path: 2D/deformable_LKA is also a package directory name
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, "2D", "deformable_LKA", "deformable_LKA.py"))
    ref = importlib.util.module_from_spec(spec)
    sys.modules[name] = ref
    spec.loader.exec_module(ref)
    return ref


class _DropPath(nn.Module):
    def __init__(self, drop_prob=0.0):
        super().__init__()
        assert drop_prob == 0.0

    def forward(self, x):
        return x


def import_reference_maxvit():
    """2D/networks/MaxViT_deform_LKA.py (DWConvLKA, Mlp, deformableLKABlock, PatchExpand, FinalPatchExpand_X4, MyDecoderLayer)."""
    _stub("fvcore"); _stub("fvcore.nn", FlopCountAnalysis=object)
    _stub("timm"); _stub("timm.models"); _stub("timm.models.layers", DropPath=_DropPath)
    p = os.path.join(REF, "2D")
    if p not in sys.path:
        sys.path.insert(0, p)
    import networks  # noqa: the reference package (namespace)
    _stub("networks.merit_lib"); _stub("networks.merit_lib.networks", MaxViT4Out_Small=object)
    import networks.MaxViT_deform_LKA as ref  # noqa
    return ref


class _MonaiConvolution(nn.Sequential):
    def __init__(self, spatial_dims, in_channels, out_channels, strides=1, kernel_size=3, act=None, norm=None, dropout=None,
                 bias=True, conv_only=False, is_transposed=False, padding=None, output_padding=None):
        super().__init__()
        assert spatial_dims == 3 and conv_only and not is_transposed and dropout is None
        self.add_module("conv", nn.Conv3d(in_channels, out_channels, kernel_size, strides, padding, bias=bias))


def _get_act_layer(name):
    kind, kw = name
    assert kind == "leakyrelu"
    return nn.LeakyReLU(**kw)


def _get_norm_layer(name, spatial_dims, channels):
    assert name == "batch" and spatial_dims == 3
    return nn.BatchNorm3d(channels)


class _Names:
    PRELU = "prelu"
    INSTANCE = "instance"


def _stub_3d_deps():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import oracle

    def deform_conv_forward(input, weight, bias, offset, kd, kh, kw, sd, sh, sw, pd, ph, pw, dd, dh, dw, group, dg, step):
        assert tuple(weight.shape[2:]) == (kd, kh, kw)
        return oracle.deform_conv3d(input, offset, weight, bias, (sd, sh, sw), (pd, ph, pw), (dd, dh, dw), group, dg)

    _stub("D3D", deform_conv_forward=deform_conv_forward)
    _stub("monai"); _stub("monai.networks"); _stub("monai.networks.blocks")
    _stub("monai.networks.blocks.convolutions", Convolution=_MonaiConvolution)
    _stub("monai.networks.layers"); _stub("monai.networks.layers.factories", Act=_Names, Norm=_Names)
    _stub("monai.networks.layers.utils", get_act_layer=_get_act_layer, get_norm_layer=_get_norm_layer)
    p = os.path.join(REF, "3D")
    if p not in sys.path:
        sys.path.insert(0, p)


def import_reference_3d(network: str = "synapse"):
    """3D/d_lka_former/network_architecture/{synapse,acdc}/transformerblock.py (LKA3d_deform, LKA_Attention3d_deform,
    TransformerBlock_3D_single_deform_LKA) and dynunet_block.UnetResBlock."""
    _stub_3d_deps()
    import importlib
    return importlib.import_module(f"d_lka_former.network_architecture.{network}.transformerblock")


def import_reference_segmentation_network():
    """3D/d_lka_former/network_architecture/neural_network.py: only its CPU-runnable static helpers
    (_compute_steps_for_sliding_window, _get_gaussian); the tiled predictor itself is CUDA-only (:299)."""
    _stub_3d_deps()
    _stub("batchgenerators"); _stub("batchgenerators.augmentations")
    _stub("batchgenerators.augmentations.utils", pad_nd_image=None)
    import importlib
    return importlib.import_module("d_lka_former.network_architecture.neural_network")

"""deformablelka_b200 -- B200 (sm_100a) native Deformable-LKA forward path.

Drop-in module surfaces of xmindflow/deformableLKA backed by hand-written CUDA behind a C ABI
(include/dlka.h).  Importing this package requires the built shared library
(``python deformablelka_b200/build.py``); there is no CPU or PyTorch fallback.
"""
from . import ops  # noqa: F401  (loads libdlka_b200.so, raises if missing)
from ._lib import LIB_PATH, launch_count  # noqa: F401
from .deform_conv3d import DeformConv as DeformConv3d, DeformConvFunction, DeformConvPack  # noqa: F401
from .deformable_LKA import DeformConv, DeformConv2d, deformable_LKA, deformable_LKA_Attention  # noqa: F401
from .lka3d import LKA3d_deform, LKA_Attention3d_deform  # noqa: F401
from .graphs import GraphedCall  # noqa: F401  (CUDA-graph replay of an inference call)
from . import acdc, sliding_window  # noqa: F401  (ACDC variant of the 3D block: acdc.LKA3d_deform, acdc.LKA_Attention3d_deform)
from .blocks import (DWConvLKA, FinalPatchExpand_X4, Mlp, MyDecoderLayer, PatchExpand,  # noqa: F401
                     TransformerBlock_3D_single_deform_LKA, deformableLKABlock)

__all__ = [
    "ops", "DeformConv", "DeformConv2d", "deformable_LKA", "deformable_LKA_Attention",
    "DeformConv3d", "DeformConvFunction", "DeformConvPack", "LKA3d_deform", "LKA_Attention3d_deform",
    "deformableLKABlock", "Mlp", "DWConvLKA", "TransformerBlock_3D_single_deform_LKA", "launch_count", "LIB_PATH", "GraphedCall",
]

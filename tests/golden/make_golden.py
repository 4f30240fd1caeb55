"""Generate golden vectors by running the UNMODIFIED reference 2D module on CPU.

Run in the build container (``/root/reference`` is not present on the GPU box):

    python tests/golden/make_golden.py

The reference file 2D/deformable_LKA/deformable_LKA.py imports ``fvcore`` at module level
(:160), which is not installed; a dummy ``fvcore.nn`` is injected so the file imports
unchanged.  The arithmetic underneath is torchvision.ops.deform_conv2d (CPU).

Outputs (small, committed): tests/golden/ref2d_*.npz holding the state_dict, input and
output of ``deformable_LKA`` and ``deformable_LKA_Attention`` for three offset regimes
(scale 0 / 1 / 8).  The 3D reference (D3D) is CUDA-only
(3D/dcn/src/deform_conv.h:46) and cannot be run here; its pins are the K1/K2/K3 identities
in tests/test_oracle.py and, on the GPU box, the compiled reference under oracle/_ref.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/2D/deformable_LKA"
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    fv = types.ModuleType("fvcore"); fvnn = types.ModuleType("fvcore.nn")
    fvnn.FlopCountAnalysis = object
    fv.nn = fvnn
    sys.modules.setdefault("fvcore", fv); sys.modules.setdefault("fvcore.nn", fvnn)
    sys.path.insert(0, REF)
    import deformable_LKA as ref  # noqa: the unmodified reference file
    return ref


def main():
    ref = import_reference()
    torch.set_num_threads(8)
    for name, cls, dim, hw in (("lka", ref.deformable_LKA, 8, (14, 12)),
                               ("attn", ref.deformable_LKA_Attention, 8, (14, 12)),
                               ("attn_c12", ref.deformable_LKA_Attention, 12, (9, 21))):
        for scale in (0.0, 1.0, 8.0):
            torch.manual_seed(1234)
            m = cls(dim).eval()
            with torch.no_grad():
                for mod_name, mod in m.named_modules():
                    if mod_name.endswith("offset_net"):
                        mod.weight.mul_(scale); mod.bias.mul_(scale)
                        if scale > 1:  # push a good share of samples out of bounds
                            mod.bias.add_(torch.randn_like(mod.bias) * scale)
            x = torch.randn(2, dim, *hw)
            with torch.no_grad():
                y = m(x)
            out = {"x": x.numpy(), "y": y.numpy()}
            for k, v in m.state_dict().items():
                out["sd." + k] = v.numpy()
            path = os.path.join(HERE, f"ref2d_{name}_s{int(scale)}.npz")
            np.savez_compressed(path, **out)
            print("wrote", path, tuple(y.shape), float(y.abs().max()))


if __name__ == "__main__":
    main()

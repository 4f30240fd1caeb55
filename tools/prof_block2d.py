"""One forward of the 2D attention block at a C2 shape (for ncu): python tools/prof_block2d.py [C] [HW] [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import deformablelka_b200 as dl

C = int(sys.argv[1]) if len(sys.argv) > 1 else 96
hw = int(sys.argv[2]) if len(sys.argv) > 2 else 56
B = int(sys.argv[3]) if len(sys.argv) > 3 else 24
torch.manual_seed(0)
with torch.no_grad():
    m = dl.deformable_LKA_Attention(C).to("cuda:0").eval()
    x = torch.randn(B, C, hw, hw, device="cuda:0")
    for _ in range(3):
        y = m(x)
    torch.cuda.synchronize()
print("ok", float(y.abs().mean()))

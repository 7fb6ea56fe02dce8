import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nndepth_b200 as nb
B, C, H, W = 8, 256, 48, 156
f1 = torch.randn(B, C, H, W, device="cuda"); f2 = torch.randn(B, C, H, W, device="cuda")
for _ in range(3):
    nb.CorrBlock1D(f1, f2, 4, 4, precision="tf32")
torch.cuda.synchronize(); print("ok")

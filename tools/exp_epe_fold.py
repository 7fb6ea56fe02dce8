"""EPE of the default bench mode against the fp32 result of the same model for several random inputs, with and
without the BatchNorm-folded fused encoder path: is the golden's 0.003 vs 0.005 px a systematic effect or the
spread of TF32 rounding noise?"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nndepth_b200.raft_stereo as rs
torch.manual_seed(0)
model = rs.BaseRAFTStereo(iters=32).eval().cuda()
model.final_only = True
orig = rs._fused_ok
for seed in (1, 2, 3, 4):
    g = torch.Generator().manual_seed(seed)
    left = (torch.rand(1, 3, 384, 1248, generator=g) * 2 - 1).cuda()
    right = (torch.rand(1, 3, 384, 1248, generator=g) * 2 - 1).cuda()
    with torch.no_grad():
        model.dense_precision = "fp32"
        rs._fused_ok = lambda x, *n: False
        ref = model(left, right)[-1]["up_disp"]
        model.dense_precision = "mixed2x"
        a = model(left, right)[-1]["up_disp"]
        rs._fused_ok = orig
        b = model(left, right)[-1]["up_disp"]
        model.dense_precision = "fp32"
        c = model(left, right)[-1]["up_disp"]
    print(f"seed {seed}: mixed2x unfolded {(a - ref).abs().mean().item():.5f}  folded {(b - ref).abs().mean().item():.5f}  "
          f"fp32 folded {(c - ref).abs().mean().item():.6f}", flush=True)

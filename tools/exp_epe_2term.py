"""Which operand of the ConvGRU convolutions must keep fp32 precision?  Variants of the tensor-core product:
  "3": [x_hi; x_lo; x_hi] x [w_hi; w_hi; w_lo]   (full 3xTF32)
  "a": [x_hi; x_lo]       x [w_hi; w_hi]         (weights rounded to TF32, activations exact)
  "b": [x_hi; x_hi]       x [w_hi; w_lo]         (activations rounded to TF32, weights exact)  <- what ships
  "1": x_hi x w_hi                                (plain TF32 with round-to-nearest operands)
EPE vs the reference golden, every other layer as in the bench mode.  Result on B200 (KITTI, 32 iterations):
3: 0.0032 px, a: 0.0113 px, b: 0.0031 px, 1: 0.0114 px -- the recurrence is sensitive to WEIGHT rounding only."""
import os, sys
import numpy as np, torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import nndepth_b200.raft_stereo as rs
from helpers import seeded_pair
g = dict(np.load(os.path.join(ROOT, "tests/golden/raft_kitti.npz")))
left, right = (t.cuda() for t in seeded_pair(g["shape"]))
ref = torch.from_numpy(g["final_up_disp"]).cuda()
torch.manual_seed(0)
model = rs.BaseRAFTStereo(iters=32).eval().cuda()
model.final_only = True
model.fuse_gru = False


def make_half_step(variant):
    def half(self, h, x, tag):
        cz, cr, cq = (getattr(self, f"conv{g_}{tag}") for g_ in "zrq")

        def conv(inp, w, b, pad):
            w = w.detach(); whi = rs.rn_tf32(w); wlo = w - whi
            ihi = rs.rn_tf32(inp); ilo = inp - ihi
            if variant == "3":
                xi, ww = torch.cat([ihi, ilo, ihi], 1), torch.cat([whi, whi, wlo], 1)
            elif variant == "a":
                xi, ww = torch.cat([ihi, ilo], 1), torch.cat([whi, whi], 1)
            elif variant == "b":
                xi, ww = torch.cat([ihi, ihi], 1), torch.cat([whi, wlo], 1)
            else:
                xi, ww = ihi, whi
            return F.conv2d(xi, ww, b, padding=pad)
        hx = torch.cat([h, x], 1)
        z = torch.sigmoid(conv(hx, cz.weight, cz.bias, cz.padding))
        r = torch.sigmoid(conv(hx, cr.weight, cr.bias, cr.padding))
        q = torch.tanh(conv(torch.cat([r * h, x], 1), cq.weight, cq.bias, cq.padding))
        return (1 - z) * h + z * q
    return half


for variant in ("3", "a", "b", "1"):
    rs.SepConvGRU._half_step_wsplit = make_half_step(variant)
    model.dense_precision = "mixed2x"
    with torch.no_grad():
        out = model(left, right)[-1]["up_disp"]
    d = (out - ref).abs()
    print(f"GRU products: {variant}  EPE={d.mean().item():.5f} px  max={d.max().item():.4f}", flush=True)

#!/usr/bin/env python
"""How smooth is the disparity field the per-iteration lookup sees?  Range of (w1 - coords) inside each group of 32
consecutive pixels of a feature row, per iteration, for the bench's noise input and the shipped KITTI pair (random-init
weights, seed 0).  Decides whether a skewed pyramid layout (lines shared by neighbouring pixels) could pay."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nndepth_b200 as nb
from nndepth_b200.raft_stereo import BaseRAFTStereo
from nndepth_b200.engine import Padder
from oracle import ref_shim

rec = []


class Rec(nb.CorrBlock1D):
    def lookup_conv1x1(self, coords, *a, **k):
        rec.append(coords.detach().clone())
        return super().lookup_conv1x1(coords, *a, **k)

    def __call__(self, coords):
        rec.append(coords.detach().clone())
        return super().__call__(coords)


torch.manual_seed(0)
model = BaseRAFTStereo(iters=32).eval().cuda()
model.dense_precision = "mixed16"
model.corr_fn = Rec
gen = torch.Generator().manual_seed(1)
noise = (torch.rand((2, 3, 375, 1242), generator=gen) * 2 - 1, torch.rand((2, 3, 375, 1242), generator=gen) * 2 - 1)
kl, kr = ref_shim.kitti_sample_pair()
for name, (l, r) in (("noise", noise), ("kitti", (kl, kr))):
    rec.clear()
    p = Padder(l.shape, 32)
    lp, rp = p.pad(l.cuda(), r.cuda())
    with torch.no_grad():
        model(lp, rp)
    out = {}
    for it in (0, 1, 4, 16, 31):
        c = rec[it]
        B, _, H, W = c.shape
        disp = torch.arange(W, device=c.device).view(1, 1, 1, W) - c
        Wg = W // 32 * 32
        g = disp[..., :Wg].reshape(B, 1, H, Wg // 32, 32)
        rng = (g.max(-1).values - g.min(-1).values).flatten()
        out[it] = {"median_range": rng.median().item(), "p90_range": rng.quantile(0.9).item(), "max_range": rng.max().item(),
                   "mean_abs_disp": disp.abs().mean().item()}
    print(name, json.dumps(out))

"""Run every kernel of the hot path once or twice at BASELINE shapes (for ncu launch lists / captures).

    python tools/run_hotpath.py [raft|igev|agcl|all] [--precision fp32|tf32] [--lookups N]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nndepth_b200 as nb  # noqa: E402


def raft(precision, lookups):
    B, C, H, W = 8, 256, 48, 156          # BASELINE configs[1]
    torch.manual_seed(0)
    f1 = torch.randn(B, C, H, W, device="cuda")
    f2 = torch.randn(B, C, H, W, device="cuda")
    coords = torch.arange(W, device="cuda").float().view(1, 1, 1, W).repeat(B, 1, H, 1) - torch.rand(B, 1, H, W, device="cuda") * 40
    for _ in range(2):
        blk = nb.CorrBlock1D(f1, f2, 4, 4, precision=precision)
    for _ in range(lookups):
        out = blk(coords)
    torch.cuda.synchronize()
    return out


def igev(batch=4):
    B, C, H, W, G = batch, 256, 120, 160, 8   # BASELINE configs[3] geometry at reduced batch
    torch.manual_seed(0)
    f1 = torch.randn(B, C, H, W, device="cuda")
    f2 = torch.randn(B, C, H, W, device="cuda")
    cv = nb.GeometryAwareCostVolume(f1, f2, [], lambda vol, feats: vol * 0.5, 4, 4, G)
    coords = torch.arange(W, device="cuda").float().view(1, 1, 1, W).repeat(B, 1, H, 1) - torch.rand(B, 1, H, W, device="cuda") * 40
    for _ in range(2):
        out = cv(coords)
    z = torch.randn(B, W, H, W, device="cuda")
    for _ in range(2):
        nb.soft_argmin(z)
    torch.cuda.synchronize()
    return out


def agcl():
    torch.manual_seed(0)
    for (H, W) in ((22, 40), (45, 80), (90, 160)):      # BASELINE configs[2] scales, N=4
        N, C = 4, 256
        f1 = torch.randn(N, C, H, W, device="cuda")
        f2 = torch.randn(N, C, H, W, device="cuda")
        flow = torch.randn(N, 2, H, W, device="cuda") * 3
        offs = torch.rand(N, 18, H, W, device="cuda") * 2 - 1
        a = nb.AGCL(f1, f2)
        for small in (False, True):
            a(flow, offs, small_patch=small, iter_mode=False)
            a(flow, None, small_patch=small, iter_mode=True)
    torch.cuda.synchronize()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="?", default="all")
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--lookups", type=int, default=4)
    args = ap.parse_args()
    if args.what in ("raft", "all"):
        raft(args.precision, args.lookups)
    if args.what in ("igev", "all"):
        igev()
    if args.what in ("agcl", "all"):
        agcl()
    print("ok")

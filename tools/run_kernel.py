"""Run one hot-path kernel a few times at its BASELINE shape (target of `ncu -k regex:<kernel>`)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nndepth_b200 as nb

what = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(0)
if what.startswith("agcl"):
    N, C, H, W = 4, 256, 90, 160
    f1 = torch.randn(N, C, H, W, device="cuda"); f2 = torch.randn(N, C, H, W, device="cuda")
    flow = torch.randn(N, 2, H, W, device="cuda") * 3
    offs = torch.rand(N, 18, H, W, device="cuda") * 2 - 1
    a = nb.AGCL(f1, f2)
    for _ in range(reps):
        if what == "agcl_offset":
            a(flow, offs, False, False)
        else:
            a(flow, None, False, True)
elif what == "softargmin":
    z = torch.randn(16, 160, 120, 160, device="cuda")
    for _ in range(reps):
        nb.soft_argmin(z)
elif what in ("igev_lookup", "igev_build", "igev_geo"):
    B, C, H, W, G = 16, 256, 120, 160, 8
    f1 = torch.randn(B, C, H, W, device="cuda"); f2 = torch.randn(B, C, H, W, device="cuda")
    cv = nb.GeometryAwareCostVolume(f1, f2, [], lambda vol, feats: vol, 4, 4, G)
    coords = torch.arange(W, device="cuda").float().view(1, 1, 1, W).repeat(B, 1, H, 1) - torch.rand(B, 1, H, W, device="cuda") * 40
    for _ in range(reps):
        if what == "igev_lookup":
            cv(coords)
        elif what == "igev_build":
            cv._build_feature_volume(f1, f2, cv._feat)
        else:
            nb.GeometryAwareCostVolume(f1, f2, [], lambda vol, feats: vol, 4, 4, G)
elif what == "squeeze":
    B, H, W, G = 16, 120, 160, 8
    from nndepth_b200.igev import InterleavedPyramid
    cv = nb.GeometryAwareCostVolume.__new__(nb.GeometryAwareCostVolume)
    torch.nn.Module.__init__(cv)
    cv.num_groups, cv.num_levels, cv.radius, cv._shape, cv._interleaved = G, 1, 4, (B, H, W, W), True
    cv._geo_il = InterleavedPyramid(B * H * W, W, 1, torch.device("cuda"))
    cv._geo_il.levels[0].normal_()
    sq = torch.nn.Conv3d(G, 1, 3, 1, 1).cuda()
    for _ in range(reps):
        cv.init_disparity(sq)
elif what in ("raft_lookup", "raft_build"):
    B, C, H, W = 8, 256, 48, 156
    f1 = torch.randn(B, C, H, W, device="cuda"); f2 = torch.randn(B, C, H, W, device="cuda")
    blk = nb.CorrBlock1D(f1, f2, 4, 4)
    coords = torch.arange(W, device="cuda").float().view(1, 1, 1, W).repeat(B, 1, H, 1) - torch.rand(B, 1, H, W, device="cuda") * 40
    for _ in range(reps):
        if what == "raft_lookup":
            blk(coords)
        else:
            nb.CorrBlock1D(f1, f2, 4, 4)
torch.cuda.synchronize()
print("ok")

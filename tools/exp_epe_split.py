"""Which dense part may run in TF32 without leaving the 0.01 px bar: feature encoder, update block, both?"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import nndepth_b200 as nb
from nndepth_b200.raft_stereo import BaseRAFTStereo
from helpers import seeded_pair

g = dict(np.load(os.path.join(ROOT, "tests/golden/raft_kitti.npz")))
torch.manual_seed(0)
model = BaseRAFTStereo(iters=32).eval().cuda()
model.final_only = True
left, right = (t.cuda() for t in seeded_pair(g["shape"]))
ref = torch.from_numpy(g["final_up_disp"]).cuda()
orig_fnet = model.forward_fnet

for enc_tf32 in (False, True):
    for upd_tf32 in (False, True):
        def forward_fnet(a, b):
            torch.backends.cudnn.allow_tf32 = enc_tf32
            out = orig_fnet(a, b)
            torch.backends.cudnn.allow_tf32 = upd_tf32
            return out
        model.forward_fnet = forward_fnet
        with torch.no_grad():
            out = model(left, right)[-1]["up_disp"]
        d = (out - ref).abs()
        print(f"encoder_tf32={enc_tf32} update_tf32={upd_tf32} EPE={d.mean().item():.5f} px  max={d.max().item():.4f}", flush=True)

"""Which sub-modules of the dense path tolerate cuDNN TF32 inside the 0.01 px bar?  Toggle allow_tf32 per module."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import nndepth_b200 as nb
from nndepth_b200.raft_stereo import BaseRAFTStereo
from helpers import seeded_pair

g = dict(np.load(os.path.join(ROOT, "tests/golden/raft_kitti.npz")))
left, right = (t.cuda() for t in seeded_pair(g["shape"]))
ref = torch.from_numpy(g["final_up_disp"]).cuda()
torch.manual_seed(0)
model = BaseRAFTStereo(iters=32).eval().cuda()
model.final_only = True
ub = model.update_block
groups = {"fnet": [model.fnet, model.cnet_proj], "motion_enc": [ub.encoder], "gru": [ub.gru], "flow_head": [ub.flow_head], "mask": [ub.mask]}
state = {"tf32": set()}
for name, mods in groups.items():
    for m in mods:
        m.register_forward_pre_hook(lambda mod, inp, name=name: setattr(torch.backends.cudnn, "allow_tf32", name in state["tf32"]))
combos = [set(), {"mask"}, {"fnet"}, {"gru"}, {"flow_head"}, {"motion_enc"}, {"mask", "fnet"}, {"mask", "fnet", "gru"},
          {"mask", "fnet", "motion_enc"}, {"mask", "fnet", "flow_head"}, {"mask", "fnet", "gru", "motion_enc", "flow_head"}]
for c in combos:
    state["tf32"] = c
    with torch.no_grad():
        out = model(left, right)[-1]["up_disp"]
    d = (out - ref).abs()
    print(f"tf32 in {sorted(c)!s:60s} EPE={d.mean().item():.5f} px  max={d.max().item():.4f}", flush=True)

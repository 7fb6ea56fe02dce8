"""Does the TF32 drift of the update block come from TRUNCATION (bias) rather than TF32 precision itself?
Round conv weights (once) and conv inputs (forward_pre_hook) to nearest TF32 and let cuDNN run TF32."""
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

def rna(x):
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)

for round_w, round_x, scope in ((False, False, "update"), (True, False, "update"), (True, True, "update"), (True, True, "all")):
    torch.manual_seed(0)
    model = BaseRAFTStereo(iters=32).eval().cuda()
    model.final_only = True
    mods = model.update_block.modules() if scope == "update" else model.modules()
    for m in mods:
        if isinstance(m, torch.nn.Conv2d):
            if round_w:
                m.weight.data = rna(m.weight.data)
            if round_x:
                m.register_forward_pre_hook(lambda mod, inp: (rna(inp[0]),))
    orig_fnet = model.forward_fnet
    def forward_fnet(a, b, orig_fnet=orig_fnet):
        torch.backends.cudnn.allow_tf32 = (scope == "all")
        out = orig_fnet(a, b)
        torch.backends.cudnn.allow_tf32 = True
        return out
    model.forward_fnet = forward_fnet
    with torch.no_grad():
        out = model(left, right)[-1]["up_disp"]
    d = (out - ref).abs()
    print(f"tf32 scope={scope} round_w={round_w} round_x={round_x} EPE={d.mean().item():.5f} px  max={d.max().item():.4f}", flush=True)

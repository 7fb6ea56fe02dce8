"""EPE of the dense precision modes (incl. the 3xTF32 ConvGRU) vs the reference golden; one pair, KITTI, 32 iters."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from nndepth_b200.raft_stereo import BaseRAFTStereo
from helpers import seeded_pair
g = dict(np.load(os.path.join(ROOT, "tests/golden/raft_kitti.npz")))
left, right = (t.cuda() for t in seeded_pair(g["shape"]))
ref = torch.from_numpy(g["final_up_disp"]).cuda()
torch.manual_seed(0)
model = BaseRAFTStereo(iters=32).eval().cuda()
model.final_only = True
for mode in ("fp32", "mixed", "mixed2x", "tf32"):
    model.dense_precision = mode
    with torch.no_grad():
        out = model(left, right)[-1]["up_disp"]
    d = (out - ref).abs()
    print(f"{mode:8s} EPE={d.mean().item():.5f} px  max={d.max().item():.4f}", flush=True)
# only the GRU in 3xTF32, everything else fp32: isolates the split's own error
model.dense_precision = None
model.update_block.gru.recurrence = "wsplit"
torch.backends.cudnn.allow_tf32 = False
with torch.no_grad():
    out = model(left, right)[-1]["up_disp"]
d = (out - ref).abs()
print(f"gru-3x only EPE={d.mean().item():.5f} px  max={d.max().item():.4f}", flush=True)

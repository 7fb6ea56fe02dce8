"""EPE of the full KITTI forward vs the reference golden for different volume precisions."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import nndepth_b200 as nb
from nndepth_b200.raft_stereo import BaseRAFTStereo
from helpers import seeded_pair, epe

g = dict(np.load(os.path.join(ROOT, "tests/golden/raft_kitti.npz")))
torch.manual_seed(0)
model = BaseRAFTStereo(iters=32).eval().cuda()
model.final_only = True
left, right = (t.cuda() for t in seeded_pair(g["shape"]))
ref = torch.from_numpy(g["final_up_disp"]).cuda()


def rna_tf32(x):
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


class RoundedTF32(nb.CorrBlock1D):
    def __init__(self, f1, f2, L=4, r=4):
        super().__init__(rna_tf32(f1.float()), rna_tf32(f2.float()), L, r, precision="tf32")


for cudnn_tf32 in (False, True):
    torch.backends.cudnn.allow_tf32 = cudnn_tf32
    for name, fn in (("fp32", lambda a, b, L, r: nb.CorrBlock1D(a, b, L, r, precision="fp32")),
                     ("tf32-kernel", lambda a, b, L, r: nb.CorrBlock1D(a, b, L, r, precision="tf32")),
                     ("tf32-rna", RoundedTF32)):
        model.corr_fn = fn
        with torch.no_grad():
            out = model(left, right)[-1]["up_disp"]
        d = (out - ref).abs()
        print(f"cudnn_tf32={cudnn_tf32} volume={name:10s} EPE={d.mean().item():.5f} px  max={d.max().item():.4f}", flush=True)

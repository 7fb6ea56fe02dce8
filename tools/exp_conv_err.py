"""Error of cuDNN's TF32 and fp16 convolutions against fp64 on the motion encoder's real inputs (which algorithm
does cuDNN pick for fp16 3x3 convolutions, and what does it cost in accuracy?)."""
import os, sys
import numpy as np, torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import nndepth_b200.raft_stereo as rs
from helpers import seeded_pair
g = dict(np.load(os.path.join(ROOT, "tests/golden/raft_kitti.npz")))
left, right = (t.cuda() for t in seeded_pair(g["shape"]))
torch.manual_seed(0)
model = rs.BaseRAFTStereo(iters=4).eval().cuda()
model.final_only = True
model.dense_precision = "mixed16"
cap = {}
orig = rs.BasicMotionEncoder.forward


def enc(self, flow, corr, cor1=None, split_flow=False):
    cap["cor1"] = cor1.detach().float().clone()
    cap["flow"] = flow.detach().clone()
    return orig(self, flow, corr, cor1=cor1, split_flow=split_flow)


rs.BasicMotionEncoder.forward = enc
with torch.no_grad():
    model(left, right)
e = model.update_block.encoder
x = cap["cor1"]
print("cor1: max", x.max().item(), "mean", x.mean().item(), "frac nonzero < 6e-5:", ((x > 0) & (x < 6e-5)).float().mean().item(),
      "frac zero:", (x == 0).float().mean().item())
for name, conv, inp in (("convc2", e.convc2, x),):
    with torch.no_grad():
        ref = F.conv2d(inp.double(), conv.weight.double(), conv.bias.double(), padding=conv.padding)
        scale = ref.abs().mean().item()
        with rs.cudnn_tf32(True):
            y = F.conv2d(inp, rs.inference_weight(conv), conv.bias, padding=conv.padding)
        print(name, "tf32 (cuDNN):      mean |err| / mean |y| =", ((y.double() - ref).abs().mean() / scale).item())
        with rs.cudnn_tf32(False):
            y = F.conv2d(inp, conv.weight, conv.bias, padding=conv.padding)
        print(name, "fp32:              ", ((y.double() - ref).abs().mean() / scale).item())
        for bench in (False, True):
            torch.backends.cudnn.benchmark = bench
            for cl in (False, True):
                xi, w = inp.half(), conv.weight.half()
                if cl:
                    xi, w = xi.contiguous(memory_format=torch.channels_last), w.contiguous(memory_format=torch.channels_last)
                y = F.conv2d(xi, w, None, padding=conv.padding)
                y32 = y.float() + conv.bias.view(1, -1, 1, 1)
                print(name, f"fp16 benchmark={bench} channels_last={cl}:", ((y32.double() - ref).abs().mean() / scale).item())
        # exact products of the fp16-rounded operands, fp32 accumulation: what a clean fp16 tensor-core kernel gives
        with rs.cudnn_tf32(False):
            y = F.conv2d(inp.half().float(), conv.weight.half().float(), conv.bias, padding=conv.padding)
        print(name, "fp16 operands, fp32 math:", ((y.double() - ref).abs().mean() / scale).item(),
              " + fp16 output:", ((y.half().double() - ref).abs().mean() / scale).item())

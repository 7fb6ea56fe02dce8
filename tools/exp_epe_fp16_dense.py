"""Can the REST of the refinement iteration (motion encoder, flow / mask heads) run as fp16 convolutions too?
fp16 weights/activations carry TF32's mantissa (cuDNN's TF32 kernels even truncate activations), the new error is
the fp16 rounding of every convolution OUTPUT.  Emulated with torch ops on top of the fp16 ConvGRU (mixed16).
  base : mixed16 as shipped (GRU fp16, other convolutions TF32)
  A    : + motion encoder fp16        B : A + heads fp16, mask logits fp16        C : B, mask logits kept fp32"""
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
model.dense_precision = "mixed16"
orig_enc, orig_ub = rs.BasicMotionEncoder.forward, rs.BasicUpdateBlock.forward


def c16(conv, x, relu=True, out32=False):
    if out32:   # fp16-rounded operands, exact fp32 accumulate and output
        with torch.backends.cudnn.flags(allow_tf32=False):
            y = F.conv2d(x.half().float(), conv.weight.half().float(), conv.bias, padding=conv.padding)
    elif BIAS32:   # fp16 convolution without bias; bias added in fp32, result rounded to fp16 again
        y = (F.conv2d(x.half(), conv.weight.half(), None, padding=conv.padding).float() + conv.bias.view(1, -1, 1, 1))
        return (F.relu(y) if relu else y).half()
    else:
        y = F.conv2d(x.half(), conv.weight.half(), conv.bias.half(), padding=conv.padding)
    return F.relu(y) if relu else y


WHICH = {"c2", "f2", "conv"}


def pick(tag, conv, x):
    if tag in WHICH:
        return c16(conv, x).float()
    if "rn" in WHICH:      # TF32 convolution on activations rounded to NEAREST tf32 (cuDNN would truncate them)
        return rs.conv_relu(conv, rs.rn_tf32(x.float()))
    return rs.conv_relu(conv, x.float())


def enc16(self, flow, corr, cor1=None, split_flow=False):
    cor = pick("c2", self.convc2, cor1)
    flo = pick("f2", self.convf2, F.relu(self.convf1(flow)))
    out = pick("conv", self.conv, torch.cat([cor, flo], 1))
    return torch.cat([out.float(), flow], 1)


def make_ub(heads16, mask32):
    def ub(self, net, inp, corr, flow, raw_mask=False, cor1=None, gru_run=None, coords=None, org=None):
        motion = self.encoder(flow, corr, cor1=cor1)
        net = gru_run.step(motion)
        if heads16:
            hidden = c16(self.mask[0], net)
            mask = c16(self.mask[2], hidden, relu=False, out32=mask32).float() - self.mask[2].bias.view(1, -1, 1, 1)
            fh = c16(self.flow_head.conv1, net)
            with torch.backends.cudnn.flags(allow_tf32=False):
                delta = self.flow_head.conv2(fh.float())
        else:
            hidden = rs.conv_relu(self.mask[0], net)
            mask = F.conv2d(hidden, rs.inference_weight(self.mask[2]), None)
            delta = self.flow_head(net)
        return net, mask, (coords + delta, coords + delta - org)
    return ub


BIAS32 = False
rs.BasicUpdateBlock.forward = make_ub(False, False)
rs.BasicMotionEncoder.forward = enc16
for which in ((), ("rn",), ("c2",), ("f2",), ("conv",), ("c2", "f2", "conv")):
    WHICH = set(which)
    with torch.no_grad():
        out = model(left, right)[-1]["up_disp"]
    d = (out - ref).abs()
    print(f"fp16 convs {which}: EPE={d.mean().item():.5f} px  max={d.max().item():.4f}", flush=True)

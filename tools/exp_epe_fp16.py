"""Would fp16 tensor-core products (2x the TF32 rate) keep the ConvGRU inside the 0.01 px bar?
fp16 has TF32's 10-bit mantissa, so RN_fp16(x) == RN_tf32(x) while |x| stays in fp16's normal range; the weights
are split [w_hi16; w_lo16].  Variants (emulated with exact fp32 convolutions on the rounded operands):
  "b"   : TF32 split (what ships)                      "h32": fp16 operands, fp32 pre-activations
  "h16" : fp16 operands, pre-activations rounded to fp16 (what cuDNN's fp16 convolution returns)
Then times cuDNN: TF32 vs fp16 convolution of the staging buffer at the bench shape."""
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
model.fuse_gru = False
stats = {}


def make_half_step(variant):
    def half(self, h, x, tag):
        cz, cr, cq = (getattr(self, f"conv{g_}{tag}") for g_ in "zrq")

        def conv(inp, w, b, pad):
            w = w.detach()
            if variant == "b":
                whi = rs.rn_tf32(w); wlo = w - whi; ihi = rs.rn_tf32(inp)
            else:
                whi = w.half().float(); wlo = (w - whi).half().float(); ihi = inp.half().float()
                stats["max_act"] = max(stats.get("max_act", 0.0), inp.abs().max().item())
            with torch.backends.cudnn.flags(allow_tf32=False):
                out = F.conv2d(torch.cat([ihi, ihi], 1), torch.cat([whi, wlo], 1), None, padding=pad)
            if variant == "h16":
                stats["max_pre"] = max(stats.get("max_pre", 0.0), out.abs().max().item())
                out = out.half().float()
            return out + b.view(1, -1, 1, 1)
        hx = torch.cat([h, x], 1)
        z = torch.sigmoid(conv(hx, cz.weight, cz.bias, cz.padding))
        r = torch.sigmoid(conv(hx, cr.weight, cr.bias, cr.padding))
        q = torch.tanh(conv(torch.cat([r * h, x], 1), cq.weight, cq.bias, cq.padding))
        return (1 - z) * h + z * q
    return half


for variant in (() if "--time-only" in sys.argv else ("b", "h32", "h16")):
    rs.SepConvGRU._half_step_wsplit = make_half_step(variant)
    model.dense_precision = "mixed2x"
    with torch.no_grad():
        out = model(left, right)[-1]["up_disp"]
    d = (out - ref).abs()
    print(f"GRU products: {variant}  EPE={d.mean().item():.5f} px  max={d.max().item():.4f}  {stats}", flush=True)

# ---- timing at the bench shape ------------------------------------------------------------------
torch.backends.cudnn.benchmark = True
cl = torch.channels_last


def timeit(fn, reps=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


for cout in (256, 128):
    for ks, pad in (((1, 5), (0, 2)), ((5, 1), (2, 0))):
        S32 = torch.randn(8, 768, 48, 156, device="cuda").contiguous(memory_format=cl)
        w32 = torch.randn(cout, 768, *ks, device="cuda").contiguous(memory_format=cl)
        S16, w16 = S32.half(), w32.half()
        with torch.no_grad(), rs.cudnn_tf32(True):
            t32 = timeit(lambda: F.conv2d(S32, w32, None, padding=pad))
            t16 = timeit(lambda: F.conv2d(S16, w16, None, padding=pad))
        gf = 2 * 8 * 48 * 156 * 768 * 5 * cout / 1e9
        print(f"conv {ks} 768->{cout}: tf32 {t32:.1f} us ({gf / t32 * 1e-3:.0f} TF/s)  fp16 {t16:.1f} us ({gf / t16 * 1e-3:.0f} TF/s)", flush=True)

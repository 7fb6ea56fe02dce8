import os, sys, statistics, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nndepth_b200 as nb
B, C, H, W = 8, 256, 48, 156
torch.manual_seed(0)
f1 = torch.randn(B, C, H, W, device="cuda"); f2 = torch.randn(B, C, H, W, device="cuda")
blk = nb.CorrBlock1D(f1, f2, 4, 4)
coords = torch.arange(W, device="cuda").float().view(1, 1, 1, W).repeat(B, 1, H, 1) - torch.rand(B, 1, H, W, device="cuda") * 40
conv = torch.nn.Conv2d(36, 256, 1).cuda()
w, b = conv.weight.detach(), conv.bias.detach()
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
def timed(fn, reps=20):
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        torch.cuda._sleep(1000000); flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)
print("lookup alone", timed(lambda: blk(coords)))
wt = blk.prepare_conv1x1_weight(w)
print("fused lookup+conv1x1+relu fp32", timed(lambda: blk.lookup_conv1x1(coords, None, b, True, weight_t=wt)))
print("fused lookup+conv1x1+relu tf32", timed(lambda: blk.lookup_conv1x1(coords, None, b, True, weight_t=wt, precision="tf32")))
print("fused tf32, channels-last fp32 out", timed(lambda: blk.lookup_conv1x1(coords, None, b, True, weight_t=wt, precision="tf32", channels_last=True)))
print("fused tf32, channels-last fp16 out", timed(lambda: blk.lookup_conv1x1(coords, None, b, True, weight_t=wt, precision="tf32", channels_last=True, half=True)))
if "--fused-only" in sys.argv:
    sys.exit(0)
with torch.no_grad():
    x = blk(coords)
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        print("cudnn conv1x1+relu alone tf32=%s" % tf32, timed(lambda: torch.relu(conv(x))))

"""Where does the host-to-host step lose time against the device-resident step?"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nndepth_b200.engine import StereoEngine
from nndepth_b200.raft_stereo import BaseRAFTStereo
torch.manual_seed(0)
m = BaseRAFTStereo(iters=32).eval(); m.dense_precision = "mixed16"
e = StereoEngine(m, device="cuda", use_cuda_graph=True)
hl = (torch.rand(8, 3, 375, 1242) * 2 - 1).pin_memory(); hr = (torch.rand(8, 3, 375, 1242) * 2 - 1).pin_memory()
dl, dr = hl.cuda(), hr.cuda()


def wall(fn, n=20, finish=None):
    for _ in range(3):
        fn()
    if finish: finish()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        fn()
    if finish: finish()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


print("device-resident infer_device      %.2f ms" % wall(lambda: e.infer_device(dl, dr)))
print("synchronous infer (H2D+fwd+D2H)    %.2f ms" % wall(lambda: e.infer(hl, hr)))
pend = []
def step():
    pend.append(e.submit(hl, hr))
    if len(pend) > 1: e.collect(pend.pop(0))
def drain():
    while pend: e.collect(pend.pop(0))
print("pipelined submit/collect           %.2f ms" % wall(step, finish=drain))
# H2D alone / D2H alone
out = torch.empty(8, 1, 375, 1242, device="cuda"); ho = torch.empty(8, 1, 375, 1242).pin_memory()
print("H2D of both images alone           %.2f ms" % wall(lambda: (dl.copy_(hl, non_blocking=True), dr.copy_(hr, non_blocking=True))))
print("D2H of the disparity alone         %.2f ms" % wall(lambda: ho.copy_(out, non_blocking=True)))
# device step while an unrelated H2D runs on another stream
cs = torch.cuda.Stream()
def both():
    with torch.cuda.stream(cs):
        dl2.copy_(hl, non_blocking=True); dr2.copy_(hr, non_blocking=True)
    e.infer_device(dl, dr)
dl2, dr2 = torch.empty_like(dl), torch.empty_like(dr)
print("infer_device + concurrent H2D      %.2f ms" % wall(both))

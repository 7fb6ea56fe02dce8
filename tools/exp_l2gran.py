"""Does cudaLimitMaxL2FetchGranularity change the lookup's DRAM over-fetch (40-byte windows in 64/128-byte granules)?"""
import ctypes, os, sys, statistics, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nndepth_b200 as nb
torch.zeros(1, device="cuda")
rt = None
for name in ("libcudart.so.12", "libcudart.so"):
    try:
        rt = ctypes.CDLL(name); break
    except OSError:
        pass
cudaLimitMaxL2FetchGranularity = 0x05
val = ctypes.c_size_t()
print("get:", rt.cudaDeviceGetLimit(ctypes.byref(val), cudaLimitMaxL2FetchGranularity), val.value)
B, C, H, W = 8, 256, 48, 156
f1 = torch.randn(B, C, H, W, device="cuda"); f2 = torch.randn(B, C, H, W, device="cuda")
blk = nb.CorrBlock1D(f1, f2, 4, 4)
coords = torch.arange(W, device="cuda").float().view(1, 1, 1, W).repeat(B, 1, H, 1) - torch.rand(B, 1, H, W, device="cuda") * 40
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
for g in (None, 32, 64, 128):
    if g is not None:
        print("set", g, "->", rt.cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, ctypes.c_size_t(g)))
        rt.cudaDeviceGetLimit(ctypes.byref(val), cudaLimitMaxL2FetchGranularity)
    print(f"granularity {val.value}: lookup {timed(lambda: blk(coords)):.1f} us")

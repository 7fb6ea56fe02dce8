"""Time the tf32 build under the NND_EXP / NND_LAG experiment knobs (kernel-only, spin kernel first)."""
import os, sys, statistics, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nndepth_b200 as nb
shapes = [(8, 256, 48, 156), (1, 256, 136, 240)]
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
exps = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "0,1,2,4,6").split(",")]
lags = (sys.argv[2] if len(sys.argv) > 2 else "2").split(",")
for (B, C, H, W) in shapes:
    f1 = torch.randn(B, C, H, W, device="cuda"); f2 = torch.randn(B, C, H, W, device="cuda")
    for lag in lags:
        os.environ["NND_LAG"] = lag
        for exp in exps:
            os.environ["NND_EXP"] = str(exp)
            ts = []
            for i in range(12):
                torch.cuda._sleep(1000000); flush.fill_(1.0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); nb.CorrBlock1D(f1, f2, 4, 4, precision="tf32"); e1.record(); e1.synchronize()
                if i >= 2: ts.append(e0.elapsed_time(e1) * 1e3)
            print(f"B{B} {H}x{W} slack={lag} exp={exp}: median {statistics.median(ts):.1f} us  min {min(ts):.1f}", flush=True)

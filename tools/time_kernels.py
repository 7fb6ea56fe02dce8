"""CUDA-event timing of the hot-path kernels at BASELINE shapes, L2 flushed before every launch."""
import argparse
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nndepth_b200 as nb  # noqa: E402

PEAK = 6545.6


def timed(fn, reps=30, flush=None):
    stream = torch.cuda.current_stream()
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        # a ~0.5 ms spin kernel first: the host enqueues everything below while the GPU is still busy, so the
        # events bracket GPU execution only (no host launch latency inside the interval)
        torch.cuda._sleep(1000000)
        if flush is not None:
            flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts), min(ts)


def report(name, us, nbytes, flops=None):
    med, best = us
    line = f"{name:44s} {med:9.1f} us (min {best:8.1f})  {nbytes / med / 1e3:8.1f} GB/s = {nbytes / med / 1e3 / PEAK * 100:5.1f}% of HBM peak"
    if flops:
        line += f"  {flops / med / 1e6:7.1f} TFLOP/s"
    print(line, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="?", default="raft")
    args = ap.parse_args()
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    if args.what in ("floor", "raft", "all"):
        # reference points for latency-bound launches: what does the same harness report for a trivial kernel and
        # for a plain copy moving the lookup's 18.4 MB?
        one = torch.zeros(32, device="cuda")
        report("floor: 1-block fill kernel", timed(lambda: one.fill_(1.0), flush=flush), 128)
        src = torch.randn(2300000, device="cuda"); dst = torch.empty_like(src)
        report("floor: copy 9.2 MB -> 9.2 MB (L2 flushed)", timed(lambda: dst.copy_(src), flush=flush), 2 * src.numel() * 4)
        report("floor: copy 9.2 MB -> 9.2 MB (L2 warm)", timed(lambda: dst.copy_(src), flush=None), 2 * src.numel() * 4)
        big = torch.randn(8 * 1024 * 1024, device="cuda"); idx = torch.randint(0, big.numel() // 16, (59904 * 4,), device="cuda") * 16
        gat = torch.empty(59904 * 4, 16, device="cuda")
        report("floor: gather 240k x 64 B rows from 32 MB (flushed)", timed(lambda: torch.index_select(big.view(-1, 16), 0, idx // 16, out=gat), flush=flush), 2 * gat.numel() * 4)
    if args.what in ("raft", "all"):
        for (B, C, H, W) in ((8, 256, 48, 156), (1, 256, 80, 160), (1, 256, 136, 240)):
            torch.manual_seed(0)
            f1 = torch.randn(B, C, H, W, device="cuda")
            f2 = torch.randn(B, C, H, W, device="cuda")
            widths = [W >> l for l in range(4)]
            build_bytes = 2 * B * C * H * W * 4 + B * H * W * sum(widths) * 4
            flops = 2 * B * H * W * W * C
            for prec in ("tf32", "fp32"):
                report(f"build {prec} B{B} {H}x{W}", timed(lambda: nb.CorrBlock1D(f1, f2, 4, 4, precision=prec), flush=flush),
                       build_bytes, flops)
            blk = nb.CorrBlock1D(f1, f2, 4, 4, precision="fp32")
            coords = torch.arange(W, device="cuda").float().view(1, 1, 1, W).repeat(B, 1, H, 1) - torch.rand(B, 1, H, W, device="cuda") * 40
            lb = B * H * W * 308
            report(f"lookup B{B} {H}x{W} (L2 flushed)", timed(lambda: blk(coords), flush=flush), lb)
            report(f"lookup B{B} {H}x{W} (L2 warm)", timed(lambda: blk(coords), flush=None), lb)
    if args.what in ("igev", "all"):
        from nndepth_b200 import _lib
        from nndepth_b200.igev import InterleavedPyramid
        B, C, H, W, G = 16, 256, 120, 160, 8
        torch.manual_seed(0)
        f1 = torch.randn(B, C, H, W, device="cuda")
        f2 = torch.randn(B, C, H, W, device="cuda")
        cv = nb.GeometryAwareCostVolume(f1, f2, [], lambda vol, feats: vol, 4, 4, G)
        vol_bytes = B * G * H * W * W * 4
        pyr_bytes = B * G * H * W * (160 + 80 + 40 + 20) * 4
        fpyr = cv._feat
        report("igev groupcorr build B16 (level 0, reference layout)",
               timed(lambda: cv._build_feature_volume(f1, f2, fpyr), reps=10, flush=flush), 2 * B * 64 * H * W * 4 + vol_bytes)
        il = InterleavedPyramid(B * H * W, W, 4, f1.device)
        report("igev interleave+pool from rows (feat) B16",
               timed(lambda: il.fill(fpyr.levels[0], 0, fpyr.pitches[0], B, W, H, W), reps=10, flush=flush), vol_bytes + pyr_bytes)
        geo = torch.randn(B, G, W, H, W, device="cuda")
        report("igev interleave+pool from (B,G,D,H,W) (geo) B16",
               timed(lambda: il.fill(geo, 1, 0, B, W, H, W), reps=10, flush=flush), vol_bytes + pyr_bytes)
        del geo, il
        coords = torch.arange(W, device="cuda").float().view(1, 1, 1, W).repeat(B, 1, H, 1) - torch.rand(B, 1, H, W, device="cuda") * 40
        report("igev dual lookup B16 (interleaved pyramids)", timed(lambda: cv(coords), reps=10, flush=flush), B * H * W * 4868)
        del cv
        z = torch.randn(B, W, H, W, device="cuda")
        report("soft-argmin B16 D160 120x160", timed(lambda: nb.soft_argmin(z), reps=10, flush=flush), z.numel() * 4 + B * H * W * 4)
    if args.what in ("agcl", "all"):
        for (H, W) in ((22, 40), (45, 80), (90, 160)):
            N, C = 4, 256
            f1 = torch.randn(N, C, H, W, device="cuda")
            f2 = torch.randn(N, C, H, W, device="cuda")
            flow = torch.randn(N, 2, H, W, device="cuda") * 3
            offs = torch.rand(N, 18, H, W, device="cuda") * 2 - 1
            a = nb.AGCL(f1, f2)
            px = N * H * W
            report(f"agcl nchw->nhwc staging (one map) N4 {H}x{W}", timed(lambda: nb.AGCL(f1, f2)._nhwc(f1), reps=10, flush=flush), 2 * px * C * 4)
            for small in (False, True):
                tag = "3x3" if small else "1x9"
                report(f"agcl offset {tag} N4 {H}x{W}", timed(lambda: a(flow, offs, small, False), reps=10, flush=flush), px * 2272)
                report(f"agcl iter   {tag} N4 {H}x{W}", timed(lambda: a(flow, None, small, True), reps=10, flush=flush), px * 2200)


if __name__ == "__main__":
    main()

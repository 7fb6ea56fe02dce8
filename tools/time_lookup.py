#!/usr/bin/env python
"""A/B timing of the fused lookup + convc1 kernel: this tree's library against the round-1 library
(tools/_old/libnndepth_b200_r1.so, mma.sync kernel), same pyramid, same coordinates, L2 flushed before every launch.

    python tools/time_lookup.py [--batches 8 64] [--reps 30]
"""
import argparse
import ctypes
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", type=int, nargs="+", default=[8, 64])
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--once", action="store_true", help="one launch per kernel and batch (for ncu)")
    ap.add_argument("--smooth", action="store_true", help="smooth disparity field instead of white noise")
    args = ap.parse_args()
    import nndepth_b200 as nb
    from nndepth_b200 import _lib
    new = _lib.load()
    libs = {"r2": new, "r2_skewed": new}
    import glob
    for old_path in sorted(glob.glob(os.path.join(ROOT, "tools", "_old", "libnndepth_b200_*.so"))):
        if args.once:
            break
        old = ctypes.CDLL(old_path)
        fn = old.nnd_corr1d_lookup_conv1x1
        fn.restype, fn.argtypes = _lib.SIGNATURES["nnd_corr1d_lookup_conv1x1"]
        libs[os.path.basename(old_path)[len("libnndepth_b200_"):-3]] = old
    dev = torch.device("cuda")
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev)
    for B in args.batches:
        C, H, W = 256, 48, 156
        torch.manual_seed(3)
        f1 = torch.randn(B, C, H, W, device=dev)
        f2 = torch.randn(B, C, H, W, device=dev)
        blk = nb.CorrBlock1D(f1, f2, 4, 4)
        del f1, f2
        base = torch.arange(W, device=dev).float().view(1, 1, 1, W).repeat(B, 1, H, 1)
        if args.smooth:
            disp = torch.nn.functional.interpolate(torch.rand(B, 1, 6, 20, device=dev) * 12, size=(H, W), mode="bilinear")
        else:
            disp = torch.rand(B, 1, H, W, device=dev) * 40
        coords = (base - disp).contiguous()
        conv = torch.nn.Conv2d(36, 256, 1).to(dev)
        wt = blk.prepare_conv1x1_weight(conv.weight.detach())
        bias = conv.bias.detach()
        px = B * H * W
        if not args.once:
            # the pass that writes the skewed copy (read the row pyramid once, write it once)
            ts = []
            Wl = [W >> l for l in range(4)]
            for _ in range(args.reps):
                blk._skew = None
                flush.fill_(1.0)
                torch.cuda._sleep(200000)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                blk.skewed_pyramid()
                e1.record(stream)
                e1.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            print(json.dumps({"B": B, "kernel": "corr1d_skew", "us_l2_flushed": statistics.median(ts),
                              "bytes_read_plus_written": 2 * 4 * px * sum(Wl)}), flush=True)
        for layout, elem in ((2, 2), (1, 4)):
            out = torch.empty(B, H, W, 256, dtype=torch.float16 if layout == 2 else torch.float32, device=dev)
            nbytes = px * (164 + 256 * elem)
            res = {}
            for name, lib in libs.items():
                def launch():
                    if name == "r2_skewed":
                        res_t = blk.lookup_conv1x1(coords, None, bias, relu=True, weight_t=wt, precision="tf32", channels_last=True,
                                                   half=layout == 2, skewed=True)
                        return
                    st = lib.nnd_corr1d_lookup_conv1x1(blk._pyr._level_ptrs, blk._pyr._width_arr, blk._pyr._pitch_arr,
                                                       _lib.ptr(coords), B, H, W, 4, 4, _lib.ptr(wt), _lib.ptr(bias), 256, 1,
                                                       _lib.PREC_TF32, layout, _lib.ptr(out), _lib.stream_ptr(coords))
                    assert st == 0, st
                if args.once:
                    launch()
                    torch.cuda.synchronize()
                    continue
                for _ in range(3):
                    launch()
                ts = []
                for _ in range(args.reps):
                    flush.fill_(1.0)
                    torch.cuda._sleep(200000)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    launch()
                    e1.record(stream)
                    e1.synchronize()
                    ts.append(e0.elapsed_time(e1) * 1e3)
                warm = []
                for _ in range(args.reps):
                    torch.cuda._sleep(200000)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    launch()
                    e1.record(stream)
                    e1.synchronize()
                    warm.append(e0.elapsed_time(e1) * 1e3)
                res[name] = {"us_l2_flushed": statistics.median(ts), "us_l2_warm": statistics.median(warm),
                             "gbs_flushed": nbytes / statistics.median(ts) / 1e3}
            if not args.once:
                print(json.dumps({"B": B, "pixels": px, "out": "fp16" if layout == 2 else "fp32", "algorithmic_bytes": nbytes,
                                  "smooth": args.smooth, **res}), flush=True)
        del blk
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

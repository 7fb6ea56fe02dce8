#!/usr/bin/env python
"""A/B timing of nnd_agcl_iter_nhwc (N4 C256 90x160) across library variants under tools/_old/libagcl_*.so."""
import ctypes, glob, json, os, statistics, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nndepth_b200 import _lib
import nndepth_b200 as nb

libs = {"tree": _lib.load()}
for path in sorted(glob.glob(os.path.join(ROOT, "tools", "_old", "libagcl_*.so"))):
    lib = ctypes.CDLL(path)
    for name in ("nnd_agcl_iter_nhwc", "nnd_agcl_offset_nhwc"):
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = _lib.SIGNATURES[name]
    libs[os.path.basename(path)[len("libagcl_"):-3]] = lib
N, C, H, W = 4, 256, 90, 160
torch.manual_seed(3)
f1, f2 = torch.randn(N, C, H, W, device="cuda"), torch.randn(N, C, H, W, device="cuda")
flow = torch.randn(N, 2, H, W, device="cuda") * 3
offs = torch.rand(N, 18, H, W, device="cuda") * 2 - 1
a = nb.AGCL(f1, f2)
l, r = a._nhwc(f1), a._nhwc(f2)
ws = torch.empty_like(l)
out = torch.empty(N, 36, H, W, device="cuda")
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
stream = torch.cuda.current_stream()
ref = None
for name, lib in libs.items():
    res = {}
    for small in (0, 1):
        def launch():
            st = lib.nnd_agcl_iter_nhwc(_lib.ptr(l), _lib.ptr(r), _lib.ptr(flow), N, C, H, W, small, _lib.ptr(ws), _lib.ptr(out), _lib.stream_ptr(l))
            assert st == 0, st
        for _ in range(3):
            launch()
        ts = []
        for _ in range(10):
            flush.fill_(1.0); torch.cuda._sleep(200000)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); launch(); e1.record(stream); e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        res["3x3" if small else "1x9"] = round(statistics.median(ts), 1)
        if small == 0:
            if ref is None:
                ref = out.clone()
            else:
                res["max_diff_vs_tree"] = (out - ref).abs().max().item()
    print(name, json.dumps(res), flush=True)

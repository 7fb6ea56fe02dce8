#!/usr/bin/env python
"""Host-side cost of one call of the per-iteration lookup through the two bindings of the same C ABI: the PyTorch operator
library (torch.ops.nndepth_b200.corr1d_lookup, what the mirror classes use) and a raw ctypes call (what round 1 used).
The problem is tiny, so the GPU is never the bottleneck: the time per call is the host's marshalling + launch."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nndepth_b200 as nb
from nndepth_b200 import _lib

B, C, H, W = 1, 32, 8, 64
f1, f2 = torch.randn(B, C, H, W, device="cuda"), torch.randn(B, C, H, W, device="cuda")
blk = nb.CorrBlock1D(f1, f2, 4, 4)
coords = torch.arange(W, device="cuda").float().view(1, 1, 1, W).repeat(B, 1, H, 1).contiguous()
lib, ops = _lib.load(), _lib.ops()
pyr = blk._pyr


def via_ops():
    return ops.corr1d_lookup(pyr.buffer, W, coords, 4, 4)


def via_ctypes():
    out = torch.empty(B, 36, H, W, dtype=torch.float32, device=coords.device)
    with torch.cuda.device(coords.device):
        _lib.check(lib.nnd_corr1d_lookup(_lib.ptr_array(pyr.levels), _lib.int_array(pyr.widths), _lib.int_array(pyr.pitches),
                                         _lib.ptr(coords), B, H, W, 4, 4, _lib.ptr(out), _lib.stream_ptr(coords)), "lookup")
    return out


def via_class():
    return blk(coords)


res = {}
for name, fn in (("torch_ops", via_ops), ("ctypes", via_ctypes), ("CorrBlock1D.__call__", via_class)):
    for _ in range(200):
        fn()
    torch.cuda.synchronize()
    n = 3000
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    res[name] = (time.perf_counter() - t0) / n * 1e6
print(json.dumps({"host_us_per_call": res, "shape": [B, C, H, W]}))

"""Bring-up check of the tcgen05 TF32 build against an fp64 contraction on the device."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nndepth_b200 as nb  # noqa: E402

shapes = [(1, 32, 2, 32, 32), (1, 64, 3, 64, 64), (2, 256, 4, 156, 156), (1, 256, 2, 160, 160), (1, 40, 2, 40, 72),
          (1, 256, 2, 240, 240), (1, 24, 1, 300, 300), (1, 16, 2, 520, 264), (8, 256, 48, 156, 156)]
if len(sys.argv) > 1:
    shapes = shapes[:int(sys.argv[1])]
for (B, C, H, W1, W2) in shapes:
    torch.manual_seed(0)
    f1 = torch.randn(B, C, H, W1, device="cuda")
    f2 = torch.randn(B, C, H, W2, device="cuda")
    L = 4 if (W2 >> 3) >= 1 else 1
    ref = torch.einsum("bchi,bchj->bhij", f1.double(), f2.double()).float() / C ** 0.5
    t0 = time.time()
    blk = nb.CorrBlock1D(f1, f2, L, 4, precision="tf32")
    torch.cuda.synchronize()
    dt = time.time() - t0
    pyr = blk.corr_pyramid
    got = pyr[0].reshape(B, H, W1, W2)
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item()
    ok_pool = True
    for l in range(L - 1):
        lo = pyr[l][:, 0]
        half = lo.shape[1] // 2
        expect = (lo[:, 0:2 * half:2] + lo[:, 1:2 * half:2]) * 0.5
        ok_pool = ok_pool and torch.equal(pyr[l + 1][:, 0], expect)
    fp32 = nb.CorrBlock1D(f1, f2, L, 4, precision="fp32").corr_pyramid[0].reshape(B, H, W1, W2)
    print(f"shape {(B, C, H, W1, W2)}: max|err|/scale = {err / scale:.2e} (fp32 path {((fp32 - ref).abs().max().item()) / scale:.2e}) "
          f"pool_exact={ok_pool} first-call {dt * 1e3:.1f} ms", flush=True)
    if not (err / scale < 2e-3 and ok_pool):
        bad = (got - ref).abs() > 2e-3 * scale
        idx = bad.nonzero()[:8].tolist()
        print("  MISMATCH at", idx, "of", int(bad.sum()), "elements; got/ref:",
              [(got[tuple(i)].item(), ref[tuple(i)].item()) for i in idx[:4]])
print("done")

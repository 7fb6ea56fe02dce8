#!/usr/bin/env python
"""Per-kernel bench of the hot path at the BASELINE.json configs that are NOT the headline (`bench.py` is
configs[1]): one JSON line per config with CUDA-event timings (L2 flushed before every launch), the
roofline of its dominant kernel and the oracle timed on the host beside it (`cpu_baseline`, bounded sample).

    python bench_hotpath.py [cfg1 cfg3 cfg4 cfg5] [--skip-cpu]

The oracle (`oracle/`) is executed here only as the CPU baseline, never on the GPU path.
"""
import json
import os
import statistics
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import nndepth_b200 as nb  # noqa: E402


def peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


PEAK, PEAK_SRC = peak()
_flush = None


def gpu_us(fn, reps=12):
    """Median microseconds of `fn` by CUDA events: a spin kernel keeps the GPU busy while the host enqueues,
    then a 256 MB fill evicts L2, then the timed launch(es)."""
    global _flush
    if _flush is None:
        _flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream()
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        torch.cuda._sleep(1000000)
        _flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


def roof(kernel, nbytes, us):
    a = nbytes / us / 1e3
    return {"kernel": kernel, "bound": "hbm", "achieved": a, "peak": PEAK, "unit": "GB/s", "frac": a / PEAK,
            "traffic": None, "peak_source": PEAK_SRC, "algorithmic_bytes_per_launch": nbytes, "us_per_launch": us}


def reference_available():
    """The unmodified reference staged under oracle/_ref (or /root/reference in the build container)."""
    from oracle import ref_shim
    if not ref_shim.available():
        return False
    ref_shim.install()
    return True


def strict_fp32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def cpu_seconds(fn, reps=2):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


def cfg1(skip_cpu):
    """RAFT-Stereo pyramid + 4-level radius-4 lookup, B1, 256 ch, 80x160 -- the reference's CPU-runnable case."""
    from oracle import torch_port
    B, C, H, W = 1, 256, 80, 160
    gen = torch.Generator().manual_seed(1)
    f1c, f2c = torch.randn(B, C, H, W, generator=gen), torch.randn(B, C, H, W, generator=gen)
    grid = torch.arange(W).float().view(1, 1, 1, W).repeat(B, 1, H, 1)
    coords_c = [grid] + [grid - torch.rand(B, 1, H, W, generator=gen) * 40 for _ in range(31)]
    f1, f2 = f1c.cuda(), f2c.cuda()
    coords = [c.cuda() for c in coords_c]
    build = gpu_us(lambda: nb.CorrBlock1D(f1, f2, 4, 4))
    blk = nb.CorrBlock1D(f1, f2, 4, 4)
    lookup = gpu_us(lambda: blk(coords[5]))
    # the same build from fp16 channels-last maps (what a cuDNN fp16 encoder hands over), read in place
    h1, h2 = (t.half().contiguous(memory_format=torch.channels_last) for t in (f1, f2))
    build16 = gpu_us(lambda: nb.CorrBlock1D(h1, h2, 4, 4))
    lookup_skew = gpu_us(lambda: blk(coords[5], skewed=True))

    def whole():
        b = nb.CorrBlock1D(f1, f2, 4, 4)
        for c in coords:
            b(c)
    total = gpu_us(whole)
    line = {"config": {"workload": "BASELINE configs[0]: pyramid build + 32 lookups, B1 C256 80x160, L4 r4"},
            "metric": "pyramid build + 32 lookups", "unit": "passes/s", "value": 1e6 / total, "dtype": "f32 (TF32 operands, RN)",
            "gpu_us": {"build": build, "lookup": lookup, "build_plus_32_lookups": total,
                       "build_from_fp16_channels_last_maps": build16, "lookup_on_skewed_copy": lookup_skew},
            "roofline": roof("corr1d_build_tf32_kernel", 2 * B * C * H * W * 4 + B * H * W * 300 * 4, build),
            "lookup_roofline": roof("corr1d_lookup_lean_kernel<9>", B * H * W * 308, lookup),
            "other_rooflines": [roof("corr1d_build_tf32_kernel<fp16 channels-last maps>",
                                     2 * B * C * H * W * 2 + B * H * W * 300 * 4, build16),
                                roof("corr1d_lookup_skewed_kernel<9>", B * H * W * 308, lookup_skew)]}
    have_ref = reference_available()
    if have_ref:
        from nndepth.models.raft_stereo.cost_volume import CorrBlock1D as RefCorr
        strict_fp32()

        def ref_gpu():
            b = RefCorr(f1, f2, 4, 4)
            for c in coords:
                b(c)
        with torch.no_grad():
            us_ref = gpu_us(ref_gpu, reps=4)
            rb = RefCorr(f1, f2, 4, 4)
            us_ref_lookup = gpu_us(lambda: rb(coords[5]), reps=4)
        line["gpu_baseline"] = {"what": "the reference's own CorrBlock1D (ATen chain) on torch.cuda, same inputs, fp32",
                                "build_plus_32_lookups_us": us_ref, "lookup_us": us_ref_lookup,
                                "speedup_whole": us_ref / total, "speedup_lookup": us_ref_lookup / lookup}
    if not skip_cpu:
        torch.set_num_threads(os.cpu_count() or 1)
        cls = RefCorr if have_ref else torch_port.CorrBlock1D

        def cpu():
            b = cls(f1c, f2c, 4, 4)
            for c in coords_c:
                b(c)
        with torch.no_grad():
            sec = cpu_seconds(cpu, reps=1)
        line["cpu_baseline"] = {"value": 1.0 / sec, "unit": "passes/s", "cores": torch.get_num_threads(),
                                "kind": "reference" if have_ref else "port",
                                "sample": "the full config (1 build + 32 lookups), " +
                                          ("the reference's CorrBlock1D as is (oracle/_ref)" if have_ref else "oracle/torch_port.py")}
    return line


def cfg3(skip_cpu):
    """CREStereo AGCL at the three cascade scales of 720x1280 (reference runs 1/32, 1/16, 1/8), N4 C256."""
    from oracle import agcl as oa
    N, C = 4, 256
    out = {"config": {"workload": "BASELINE configs[2]: AGCL, N4 C256 at 22x40 / 45x80 / 90x160, offset + iter mode, 1x9 + 3x3"},
           "metric": "AGCL calls", "unit": "calls/s", "dtype": "f32", "gpu_us": {}}
    for (H, W) in ((22, 40), (45, 80), (90, 160)):
        torch.manual_seed(3)
        f1, f2 = torch.randn(N, C, H, W, device="cuda"), torch.randn(N, C, H, W, device="cuda")
        flow = torch.randn(N, 2, H, W, device="cuda") * 3
        offs = torch.rand(N, 18, H, W, device="cuda") * 2 - 1
        a = nb.AGCL(f1, f2)
        for small in (False, True):
            tag = "3x3" if small else "1x9"
            out["gpu_us"][f"offset_{tag}_{H}x{W}"] = gpu_us(lambda: a(flow, offs, small, False), reps=8)
            out["gpu_us"][f"iter_{tag}_{H}x{W}"] = gpu_us(lambda: a(flow, None, small, True), reps=8)
        out["gpu_us"][f"staging_nhwc_one_map_{H}x{W}"] = gpu_us(lambda: nb.AGCL(f1, f2)._nhwc(f1), reps=8)
        if (H, W) == (90, 160):
            # a disparity-like field (smooth in x and y) instead of white noise: neighbouring pixels then
            # gather neighbouring corner vectors, which is what the cascade actually feeds the layer
            yy, xx = torch.meshgrid(torch.arange(H, device="cuda").float(), torch.arange(W, device="cuda").float(), indexing="ij")
            smooth = torch.stack([-(8 + 6 * torch.sin(xx / 23) * torch.cos(yy / 17)), 0.3 * torch.sin(yy / 9)], 0)
            smooth = smooth[None].repeat(N, 1, 1, 1).contiguous()
            out["gpu_us"]["offset_1x9_90x160_smooth_flow"] = gpu_us(lambda: a(smooth, offs, False, False), reps=8)
            out["gpu_us"]["iter_1x9_90x160_smooth_flow"] = gpu_us(lambda: a(smooth, None, False, True), reps=8)
    us = out["gpu_us"]["offset_1x9_90x160"]
    out["value"] = 1e6 / us
    out["roofline"] = roof("agcl_cl_kernel<0> (offset mode, 90x160)", N * 90 * 160 * 2272, us)
    out["roofline"]["note"] = ("gather-bound, not HBM-bound: 36 KB of corner vectors are gathered per pixel (2.2 GB through "
                               "L1, 1.6 GB from L2) for 2.3 KB of compulsory traffic; DRAM moves only 123 MB (ncu)")
    out["other_rooflines"] = [roof("agcl_iter_fused_kernel<1x9> (iter mode, 90x160)", N * 90 * 160 * 2200, out["gpu_us"]["iter_1x9_90x160"]),
                              roof("agcl_iter_fused_kernel<3x3> (iter mode, 90x160)", N * 90 * 160 * 2200, out["gpu_us"]["iter_3x3_90x160"]),
                              roof("agcl_cl4_kernel<0> (offset mode 3x3, 90x160)", N * 90 * 160 * 2272, out["gpu_us"]["offset_3x3_90x160"])]
    have_ref = reference_available()
    if have_ref:
        from nndepth.models.cre_stereo.cost_volume import AGCL as RefAGCL
        strict_fp32()
        H, W = 90, 160
        torch.manual_seed(3)
        f1, f2 = torch.randn(N, C, H, W, device="cuda"), torch.randn(N, C, H, W, device="cuda")
        flow = torch.randn(N, 2, H, W, device="cuda") * 3
        offs = torch.rand(N, 18, H, W, device="cuda") * 2 - 1
        ra = RefAGCL(f1, f2)
        with torch.no_grad():
            ref_us = {"offset_1x9_90x160": gpu_us(lambda: ra(flow, offs, False, False), reps=3),
                      "offset_3x3_90x160": gpu_us(lambda: ra(flow, offs, True, False), reps=3),
                      "iter_1x9_90x160": gpu_us(lambda: ra(flow, None, False, True), reps=3),
                      "iter_3x3_90x160": gpu_us(lambda: ra(flow, None, True, True), reps=3)}
        out["gpu_baseline"] = {"what": "the reference's own AGCL (ATen chain) on torch.cuda, same inputs, fp32", "us": ref_us,
                               "speedup": {k: v / out["gpu_us"][k] for k, v in ref_us.items()}}
        if not skip_cpu:
            torch.set_num_threads(os.cpu_count() or 1)
            rc = RefAGCL(f1.cpu(), f2.cpu())
            fc, oc = flow.cpu(), offs.cpu()
            with torch.no_grad():
                sec = cpu_seconds(lambda: rc(fc, oc, False, False), reps=1)
            out["cpu_baseline"] = {"value": 1.0 / sec, "unit": "calls/s", "cores": torch.get_num_threads(), "kind": "reference",
                                   "sample": "offset mode 1x9 at the full config (N4 C256 90x160), the reference's AGCL as is"}
    elif not skip_cpu:
        rng = np.random.default_rng(3)
        H, W = 22, 40
        f1, f2 = rng.standard_normal((1, C, H, W), dtype=np.float32), rng.standard_normal((1, C, H, W), dtype=np.float32)
        flow = (rng.standard_normal((1, 2, H, W)) * 3).astype(np.float32)
        offs = rng.uniform(-1, 1, (1, 18, H, W)).astype(np.float32)
        sec = cpu_seconds(lambda: oa.corr_att_offset(f1, f2, flow, offs, False), reps=1)
        out["cpu_baseline"] = {"value": 1.0 / (sec * 4 * 90 * 160 / (H * W)), "unit": "calls/s", "cores": 1, "kind": "port",
                               "sample": f"offset 1x9 on 1 x 256 x {H}x{W} ({sec:.2f} s, numpy oracle), scaled by pixels to N4 90x160"}
    return out


def cfg4(skip_cpu):
    """IGEV-Stereo: G8 group-wise volume, interleaved pyramids, dual lookup, soft-argmin at 480x640, B16."""
    from oracle import igev as oi
    from nndepth_b200.igev import InterleavedPyramid
    B, C, H, W, G = 16, 256, 120, 160, 8
    torch.manual_seed(0)
    f1, f2 = torch.randn(B, C, H, W, device="cuda"), torch.randn(B, C, H, W, device="cuda")
    cv = nb.GeometryAwareCostVolume(f1, f2, [], lambda vol, feats: vol, 4, 4, G)
    vol_bytes = B * G * H * W * W * 4
    pyr_bytes = B * G * H * W * 300 * 4
    us = {}
    us["groupcorr_build_level0"] = gpu_us(lambda: cv._build_feature_volume(f1, f2, cv._feat), reps=6)
    il = InterleavedPyramid(B * H * W, W, 4, f1.device)
    us["interleave_pool_feat"] = gpu_us(lambda: il.fill(cv._feat.levels[0], 0, cv._feat.pitches[0], B, W, H, W), reps=6)
    geo = torch.randn(B, G, W, H, W, device="cuda")
    us["interleave_pool_geo"] = gpu_us(lambda: il.fill(geo, 1, 0, B, W, H, W), reps=6)
    del geo, il
    coords = torch.arange(W, device="cuda").float().view(1, 1, 1, W).repeat(B, 1, H, 1) - torch.rand(B, 1, H, W, device="cuda") * 40
    us["dual_lookup"] = gpu_us(lambda: cv(coords), reps=6)
    # initial disparity: cv_squeezer Conv3d(8->1) + softmax + regress_disparity (igev_stereo/model.py:143-146)
    torch.manual_seed(1)
    squeezer = torch.nn.Conv3d(G, 1, 3, 1, 1).cuda()
    with torch.no_grad():
        us["squeeze_soft_argmin_fused"] = gpu_us(lambda: cv.init_disparity(squeezer), reps=6)

        def unfused():
            g0 = cv._geo_il.reference_level(0, B, H, W).reshape(B, G, H, W, W).permute(0, 1, 4, 2, 3)
            return nb.soft_argmin(squeezer(g0).squeeze(1))

        def torch_chain(g0):
            p = torch.softmax(squeezer(g0).squeeze(1), dim=1)
            return -(torch.arange(W, device="cuda").float().view(1, -1, 1, 1) * p).sum(1, keepdim=True)

        g0 = cv._geo_il.reference_level(0, B, H, W).reshape(B, G, H, W, W).permute(0, 1, 4, 2, 3)
        us["squeeze_cudnn_conv3d_tf32_plus_torch_softmax"] = gpu_us(lambda: torch_chain(g0), reps=3)
        with torch.backends.cudnn.flags(allow_tf32=False):
            us["squeeze_cudnn_conv3d_fp32_plus_torch_softmax"] = gpu_us(lambda: torch_chain(g0), reps=3)
            diff = (cv.init_disparity(squeezer) - torch_chain(g0)).abs().max().item()
        del g0
    del cv
    z = torch.randn(B, W, H, W, device="cuda")
    us["soft_argmin"] = gpu_us(lambda: nb.soft_argmin(z), reps=6)
    del z
    out = {"config": {"workload": "BASELINE configs[3]: IGEV G8 all-pairs volume (D=160) + geometry volume + soft-argmin, 480x640, B16"},
           "metric": "IGEV dual lookups", "unit": "lookups/s", "value": 1e6 / us["dual_lookup"], "dtype": "f32", "gpu_us": us,
           "roofline": roof("gev_lookup_kernel", B * H * W * 4868, us["dual_lookup"]),
           "other_rooflines": [roof("groupcorr_build_kernel<8>", 2 * B * 64 * H * W * 4 + vol_bytes, us["groupcorr_build_level0"]),
                               roof("gev_interleave_dmajor_kernel", vol_bytes + pyr_bytes, us["interleave_pool_feat"]),
                               roof("gev_interleave_wmajor_kernel", vol_bytes + pyr_bytes, us["interleave_pool_geo"]),
                               roof("soft_argmin_kernel<4>", B * W * H * W * 4 + B * H * W * 4, us["soft_argmin"]),
                               dict(roof("gev_squeeze_soft_argmin_kernel", vol_bytes + B * H * W * 4, us["squeeze_soft_argmin_fused"]),
                                    fp32_gflop=2 * 216 * B * W * H * W / 1e9,
                                    ffma_tflops=2 * 216 * B * W * H * W / us["squeeze_soft_argmin_fused"] / 1e6,
                                    max_abs_diff_vs_torch_fp32_px=diff)]}
    have_ref = reference_available()
    if have_ref:
        from nndepth.models.igev_stereo.cost_volume import GeometryAwareCostVolume as RefGEV
        strict_fp32()
        torch.manual_seed(0)
        f1, f2 = torch.randn(B, C, H, W, device="cuda"), torch.randn(B, C, H, W, device="cuda")
        with torch.no_grad():
            rcv = RefGEV(f1, f2, [], lambda vol, feats: vol, 4, 4, G)
            ref_lookup = gpu_us(lambda: rcv(coords), reps=3)
            del rcv
            torch.cuda.empty_cache()
            z = torch.randn(B, W, H, W, device="cuda")
            disp_values = torch.arange(W, device="cuda").float().view(1, -1, 1, 1)
            ref_soft = gpu_us(lambda: -torch.sum(disp_values * torch.softmax(z, dim=1), dim=1, keepdim=True), reps=3)
            del z
        out["gpu_baseline"] = {"what": "the reference's own GeometryAwareCostVolume.forward / softmax + regress_disparity "
                                       "(ATen chains) on torch.cuda, same shapes, fp32",
                               "dual_lookup_us": ref_lookup, "soft_argmin_us": ref_soft,
                               "speedup_lookup": ref_lookup / us["dual_lookup"], "speedup_soft_argmin": ref_soft / us["soft_argmin"]}
        if not skip_cpu:
            torch.set_num_threads(os.cpu_count() or 1)
            bs = 4                                          # bounded sample: a quarter of the batch (the CPU build alone is ~10 s at 16)
            with torch.no_grad():
                rc = RefGEV(f1[:bs].cpu(), f2[:bs].cpu(), [], lambda vol, feats: vol, 4, 4, G)
                cc = coords[:bs].cpu()
                sec = cpu_seconds(lambda: rc(cc), reps=1)
            out["cpu_baseline"] = {"value": 1.0 / (sec * B / bs), "unit": "lookups/s", "cores": torch.get_num_threads(),
                                   "kind": "reference",
                                   "sample": f"the reference's dual lookup on {bs} of the 16 pairs ({sec:.2f} s), scaled x{B // bs}"}
        del f1, f2
    elif not skip_cpu:
        rng = np.random.default_rng(0)
        b, h = 1, 8
        g1, g2 = rng.standard_normal((b, C, h, W), dtype=np.float32), rng.standard_normal((b, C, h, W), dtype=np.float32)
        vol = oi.groupwise_volume(g1, g2, G)
        fp, gp = oi.volume_pyramids(vol, vol.transpose(0, 1, 4, 2, 3), 4)
        cc = (np.broadcast_to(np.arange(W, dtype=np.float32), (b, 1, h, W)) - rng.uniform(0, 40, (b, 1, h, W))).astype(np.float32)
        sec = cpu_seconds(lambda: oi.gev_lookup(fp, gp, cc, 4, 4, G), reps=1)
        out["cpu_baseline"] = {"value": 1.0 / (sec * B * H / (b * h)), "unit": "lookups/s", "cores": 1, "kind": "port",
                               "sample": f"dual lookup on {b} x {h} rows x {W} ({sec:.2f} s, numpy oracle), scaled by pixels to B16 120x160"}
    return out


def cfg5(skip_cpu):
    """RAFT-Stereo 1080x1920 single pair: features 136x240; one of 8 row bands (17 rows) and the whole map on one GPU."""
    B, C, H, W = 1, 256, 136, 240
    torch.manual_seed(5)
    f1, f2 = torch.randn(B, C, H, W, device="cuda"), torch.randn(B, C, H, W, device="cuda")
    coords = torch.arange(W, device="cuda").float().view(1, 1, 1, W).repeat(B, 1, H, 1) - torch.rand(B, 1, H, W, device="cuda") * 60
    us = {}
    us["build_full"] = gpu_us(lambda: nb.CorrBlock1D(f1, f2, 4, 4))
    blk = nb.CorrBlock1D(f1, f2, 4, 4)
    us["lookup_full"] = gpu_us(lambda: blk(coords))
    b1, b2, bc = f1[:, :, :17].contiguous(), f2[:, :, :17].contiguous(), coords[:, :, :17].contiguous()
    us["build_band17"] = gpu_us(lambda: nb.CorrBlock1D(b1, b2, 4, 4))
    band = nb.CorrBlock1D(b1, b2, 4, 4)
    us["lookup_band17"] = gpu_us(lambda: band(bc))
    nbytes = 2 * B * C * H * W * 4 + B * H * W * 450 * 4
    line = {"config": {"workload": "BASELINE configs[4]: RAFT-Stereo 1080x1920 pair, features 136x240, 8 row bands of 17 rows"},
            "metric": "pyramid build", "unit": "builds/s", "value": 1e6 / us["build_full"], "dtype": "f32 (TF32 operands, RN)",
            "gpu_us": us, "roofline": roof("corr1d_build_tf32_kernel", nbytes, us["build_full"]),
            "lookup_roofline": roof("corr1d_lookup_lean_kernel<9>", B * H * W * 308, us["lookup_full"]),
            "note": "a 17-row band is 17 row jobs on 148 SMs: the sharded run is launch-latency bound per GPU"}
    if reference_available():
        from nndepth.models.raft_stereo.cost_volume import CorrBlock1D as RefCorr
        strict_fp32()
        with torch.no_grad():
            ref_build = gpu_us(lambda: RefCorr(f1, f2, 4, 4), reps=4)
            rb = RefCorr(f1, f2, 4, 4)
            ref_lookup = gpu_us(lambda: rb(coords), reps=4)
        line["gpu_baseline"] = {"what": "the reference's own CorrBlock1D (ATen chain) on torch.cuda, same inputs, fp32",
                                "build_us": ref_build, "lookup_us": ref_lookup,
                                "speedup_build": ref_build / us["build_full"], "speedup_lookup": ref_lookup / us["lookup_full"]}
        if not skip_cpu:
            torch.set_num_threads(os.cpu_count() or 1)
            c1, c2, cc = f1.cpu(), f2.cpu(), coords.cpu()

            def cpu():
                b = RefCorr(c1, c2, 4, 4)
                for _ in range(4):
                    b(cc)
            with torch.no_grad():
                sec = cpu_seconds(cpu, reps=1)
            line["cpu_baseline"] = {"value": 1.0 / sec, "unit": "passes/s", "cores": torch.get_num_threads(), "kind": "reference",
                                    "sample": "1 build + 4 lookups of the full 136x240 map, the reference's CorrBlock1D as is"}
    return line


def cfg5_sharded(standalone=True):
    """BASELINE configs[4] as it is meant to run: ONE 1080x1920 pair, its 136 epipolar feature rows split into
    row bands across the ranks of a torchrun job (one process per GPU).  Every rank builds the pyramid of its
    band and runs the 32 lookups on it -- no halo, no collective in the data path -- and one NCCL all-gather per
    forward collects the final lookup (the disparity-map gather of the full model).  Timed on the device, max
    over ranks; rank 0 prints the line."""
    import torch.distributed as dist
    from nndepth_b200.engine import gather_row_bands, row_band
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1 and standalone:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    B, C, H, W = 1, 256, 136, 240
    gen = torch.Generator().manual_seed(5)           # every rank draws the same pair, then keeps its band
    f1 = torch.randn(B, C, H, W, generator=gen)
    f2 = torch.randn(B, C, H, W, generator=gen)
    coords = [torch.arange(W).float().view(1, 1, 1, W).repeat(B, 1, H, 1) - torch.rand(B, 1, H, W, generator=gen) * 60 for _ in range(32)]
    b1, (h0, h1) = row_band(f1, rank, world)
    b2, _ = row_band(f2, rank, world)
    b1, b2 = b1.to(device), b2.to(device)
    bc = [row_band(c, rank, world)[0].to(device) for c in coords]

    def forward():
        blk = nb.CorrBlock1D(b1, b2, 4, 4)
        out = None
        for c in bc:
            out = blk(c)
        return gather_row_bands(out, H, world) if world > 1 else out

    for _ in range(3):
        full = forward()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    steps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        full = forward()
    e1.record()
    torch.cuda.synchronize(device)
    eager_ms = e0.elapsed_time(e1) / steps

    # the same band work (build + 32 lookups) as ONE CUDA graph: 33 launches of <= 2 MB each are host-launch bound when
    # issued eagerly; replayed from a graph the band runs at device speed, and the gather follows on the stream
    side = torch.cuda.Stream(device=device)
    side.wait_stream(torch.cuda.current_stream(device))
    with torch.cuda.stream(side):
        band_out = None
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            blk = nb.CorrBlock1D(b1, b2, 4, 4)
            for c in bc:
                band_out = blk(c)
    torch.cuda.current_stream(device).wait_stream(side)

    def forward_graphed():
        graph.replay()
        return gather_row_bands(band_out, H, world) if world > 1 else band_out

    for _ in range(3):
        full = forward_graphed()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        full = forward_graphed()
    e1.record()
    torch.cuda.synchronize(device)
    graph_ms = e0.elapsed_time(e1) / steps

    # SURVEY 8(e), second option: the update block runs REPLICATED on every rank, so each iteration's band lookup
    # (1, 36, h, 240) is all-gathered right away (32 collectives per forward, 0.6 MB each) instead of one gather at the end
    def forward_gather_every_iteration():
        blk = nb.CorrBlock1D(b1, b2, 4, 4)
        out = None
        for c in bc:
            out = blk(c)
            out = gather_row_bands(out, H, world) if world > 1 else out
        return out

    for _ in range(2):
        every = forward_gather_every_iteration()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        every = forward_gather_every_iteration()
    e1.record()
    torch.cuda.synchronize(device)
    ms = torch.tensor([graph_ms, eager_ms, e0.elapsed_time(e1) / steps], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    eager_ms, every_ms = ms[1].item(), ms[2].item()
    same_every = bool(torch.equal(every, full))
    ms = ms[:1]
    ok, line = True, None
    if rank == 0:
        # the gathered lookup equals the unsharded one bit for bit (rows are independent)
        whole = nb.CorrBlock1D(f1.to(device), f2.to(device), 4, 4)(coords[-1].to(device))
        ok = bool(torch.equal(full, whole))
        line = {"name": "cfg5_sharded", "data": "synthetic", "n_gpus": world, "scaling": "strong",
                "config": {"workload": "BASELINE configs[4]: RAFT-Stereo 1080x1920 single pair, features 136x240, "
                                       f"row-band-sharded x{world} (bands of {h1 - h0} rows on rank 0), build + 32 lookups + 1 gather"},
                "metric": "pyramid build + 32 lookups of one 1080x1920 pair", "unit": "pairs/s", "value": 1e3 / ms.item(),
                "ms_per_step": ms.item(), "ms_per_step_eager": eager_ms, "cuda_graph": True,
                "ms_per_step_gather_every_iteration": every_ms, "gather_every_iteration_equals_final_gather": same_every,
                "dtype": "f32 (TF32 operands, RN)", "gathered_equals_unsharded": ok,
                "note": "the band's build + 32 lookups are replayed as one CUDA graph (eager, the 33 launches of 0.3-2 MB each "
                        "are host-launch bound: ms_per_step_eager); the all-gather of the last lookup follows on the stream; "
                        "gather_every_iteration = SURVEY 8(e)'s replicated-GRU variant (32 collectives per forward)"}
        if standalone:
            print(json.dumps(line), flush=True)
    if world > 1 and standalone:
        dist.barrier()
        dist.destroy_process_group()
    if not ok:
        raise SystemExit("row-band gather differs from the unsharded lookup")
    return line


def main():
    if "cfg5_sharded" in sys.argv[1:]:
        nb.load_library()
        return cfg5_sharded()
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    skip_cpu = "--skip-cpu" in sys.argv
    which = args or ["cfg1", "cfg3", "cfg4", "cfg5"]
    if not torch.cuda.is_available():
        raise SystemExit("bench_hotpath.py needs a CUDA device: nndepth_b200 has no CPU fallback")
    nb.load_library()
    table = {"cfg1": cfg1, "cfg3": cfg3, "cfg4": cfg4, "cfg5": cfg5}
    for name in which:
        line = table[name](skip_cpu)
        line["name"] = name
        line["data"] = "synthetic"
        print(json.dumps(line), flush=True)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

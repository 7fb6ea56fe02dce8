#!/usr/bin/env python
"""Headline benchmark: RAFT-Stereo inference on KITTI-sized pairs (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this repo's B200 path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU

One "step" = one full stereo forward (feature encoder, correlation pyramid build, 32 GRU iterations
each with a fused 4-level radius-4 lookup, convex upsampling of every iteration as the reference's
``forward`` does) over a batch of 8 synthetic 375x1242 pairs per GPU (padded to 384x1248 like the
reference's ``evaluate.py``), random-init weights of the reference architecture (seed 0).

Prints ONE JSON line (rank 0).  ``value`` = pairs/s with inputs resident in HBM; ``e2e`` = pairs/s
through ``StereoEngine.submit/collect`` with pinned HOST buffers (H2D + D2H inside the timed region);
``roofline`` = the per-iteration lookup kernel against measured HBM bandwidth; ``parity`` = final EPE of
the timed configuration against the reference model, measured in the run; ``value_fp32`` = the same step
with every dense layer in fp32; ``gpu_baseline`` = the unmodified reference on torch.cuda; ``cpu_baseline`` =
the unmodified reference (oracle/_ref) on all host threads, bounded sample; ``hotpath`` = BASELINE configs
0 / 2 / 3 / 4 kernel by kernel; ``strong`` = the same 8 pairs split across the GPUs.  Multi-GPU: one process
per GPU (torchrun), batch-sharded, no data-path collective; the only NCCL call gathers the output maps
(on its own stream, overlapped with the next step).
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PAIRS_PER_GPU = 8
IMAGE_HW = (375, 1242)
ITERS = 32
METRIC = "RAFT-Stereo KITTI 375x1242 32-iter inference throughput"
UNIT = "pairs/s"
LOOKUP_BYTES_PER_PIXEL = 308      # 4*(2r+2)*4 window + 4 coords + 36*4 out


def workload_config(n_gpus):
    return {
        "workload": "BASELINE configs[1]: RAFT-Stereo full inference, KITTI 375x1242 (padded 384x1248), "
                    "32 GRU iterations, batch 8 synthetic pairs per GPU",
        "pairs_per_gpu": PAIRS_PER_GPU, "global_pairs": PAIRS_PER_GPU * n_gpus, "image": list(IMAGE_HW),
        "iters": ITERS, "weights": "random init (seed 0) of the reference BaseRAFTStereo architecture",
        "parallelism": f"batch-sharded x{n_gpus}, no data-path collective, one all_gather of the disparity maps",
        "l2": "inputs larger than L2: one GRU iteration streams > 2 GB of activations (126 MB L2), so the 70 MB pyramid "
              "and every tensor are cold again at the next iteration; isolated kernel timings flush L2 (256 MB write) "
              "before every timed launch",
    }


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    REASONS = {0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def run(self):
        if self._nv is None:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self._nv.nvmlDeviceGetClockInfo(self._h, self._nv.NVML_CLOCK_SM))
                mask = self._nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ------------------------------------------------------------------------------------------------
# reference legs: the UNMODIFIED reference (oracle/_ref, staged by __graft_entry__.build()) on the host CPU and on
# torch.cuda; the oracle port only when the staged copy is missing.  The only places bench.py executes oracle/.
# ------------------------------------------------------------------------------------------------
def reference_model(seed=0):
    """(model, Padder class, kind): the reference's own BaseRAFTStereo (stock code path, kind "reference"), else the
    repo's model shell with the oracle's torch-op correlation (kind "port")."""
    from oracle import ref_shim
    torch.manual_seed(seed)
    if ref_shim.available():
        ref_shim.install()
        from nndepth.models.raft_stereo.model import BaseRAFTStereo as RefModel
        from nndepth.data.dataloaders.utils import Padder as RefPadder
        return RefModel(iters=ITERS).eval(), (lambda shape: RefPadder(shape[-2:], divis_by=32)), "reference"
    from nndepth_b200.raft_stereo import BaseRAFTStereo
    from nndepth_b200.engine import Padder
    from oracle import torch_port
    model = BaseRAFTStereo(iters=ITERS).eval()
    model.corr_fn = torch_port.CorrBlock1D
    return model, (lambda shape: Padder(shape, 32)), "port"


def cpu_forward_seconds(steps, warmup, pairs=1):
    """Seconds per step of the reference on the host cores: pad -> forward -> unpad (evaluate.py:146-156)."""
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    model, make_padder, kind = reference_model()
    gen = torch.Generator().manual_seed(1)
    left = torch.rand((pairs, 3) + IMAGE_HW, generator=gen) * 2 - 1
    right = torch.rand((pairs, 3) + IMAGE_HW, generator=gen) * 2 - 1
    padder = make_padder(left.shape)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            lp, rp = padder.pad(left, right)
            out = padder.unpad(model(lp, rp)[-1]["up_disp"])
            float(out.sum())
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return sum(times) / len(times), torch.get_num_threads(), kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sec, threads, kind = cpu_forward_seconds(args.steps, args.warmup, pairs=1)
    value = 1.0 / sec
    sample = (f"1 pair of {IMAGE_HW[0]}x{IMAGE_HW[1]} per step ({ITERS} iterations), {args.steps} timed steps, "
              + ("the reference's BaseRAFTStereo as is (oracle/_ref)" if kind == "reference"
                 else "model shell in torch CPU ops + oracle/torch_port.py correlation"))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def reference_on_cuda(device, left, right, steps=3):
    """The reference model on torch.cuda, same weights (seed 0) and inputs: (disparity in strict fp32, ms per forward in
    strict fp32, ms per forward with PyTorch's stock flags).  The like-for-like GPU baseline and the parity reference."""
    model, make_padder, kind = reference_model()
    model = model.to(device)
    padder = make_padder(left.shape)
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    out, ms = None, {}
    try:
        for label, tf32 in (("stock_flags", None), ("strict_fp32", False)):
            torch.backends.cudnn.benchmark = False
            if tf32 is None:
                torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = True, False   # PyTorch defaults
            else:
                torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = tf32
            with torch.no_grad():
                for i in range(steps + 1):
                    if i == 1:
                        torch.cuda.synchronize(device)
                        t0 = time.perf_counter()
                    lp, rp = padder.pad(left, right)
                    out = padder.unpad(model(lp, rp)[-1]["up_disp"])
                torch.cuda.synchronize(device)
            ms[label] = (time.perf_counter() - t0) / steps * 1e3
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = saved
    return out, ms, kind


# ------------------------------------------------------------------------------------------------
# isolated lookup kernel timing (roofline leg)
# ------------------------------------------------------------------------------------------------
def time_lookup_kernel(device, reps=40):
    import nndepth_b200 as nb
    B, C, H, W = PAIRS_PER_GPU, 256, 48, 156
    torch.manual_seed(3)
    f1 = torch.randn(B, C, H, W, device=device)
    f2 = torch.randn(B, C, H, W, device=device)
    blk = nb.CorrBlock1D(f1, f2, 4, 4)
    coords = (torch.arange(W, device=device).float().view(1, 1, 1, W).repeat(B, 1, H, 1)
              - torch.rand(B, 1, H, W, device=device) * 40)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=device)   # 256 MB > L2
    stream = torch.cuda.current_stream(device)
    for _ in range(5):
        blk(coords)
    cold, warm = [], []
    for _ in range(reps):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        blk(coords)
        e1.record(stream)
        e1.synchronize()
        cold.append(e0.elapsed_time(e1))
    for _ in range(reps):
        # keep the GPU busy (~100 us spin) while the host enqueues, so the events bracket the kernel only
        torch.cuda._sleep(200000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        blk(coords)
        e1.record(stream)
        e1.synchronize()
        warm.append(e0.elapsed_time(e1))
    build = []
    for _ in range(10):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        nb.CorrBlock1D(f1, f2, 4, 4)
        e1.record(stream)
        e1.synchronize()
        build.append(e0.elapsed_time(e1))
    # the form the fp16 step runs: fp16 channels-last maps straight from the encoder, K-major fp16 operands
    h1 = f1.half().contiguous(memory_format=torch.channels_last)
    h2 = f2.half().contiguous(memory_format=torch.channels_last)
    build16 = []
    for _ in range(13):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        nb.CorrBlock1D(h1, h2, 4, 4)
        e1.record(stream)
        e1.synchronize()
        build16.append(e0.elapsed_time(e1))
    n_pix = B * H * W
    return {"lookup_ms_l2_flushed": statistics.median(cold), "lookup_ms_l2_warm": statistics.median(warm),
            "build_ms_l2_flushed": statistics.median(build), "pixels": n_pix,
            "build_f16_nhwc_ms_l2_flushed": statistics.median(build16[3:]),
            "build_f16_nhwc_bytes": 2 * B * C * H * W * 2 + B * H * W * (156 + 78 + 39 + 19) * 4,
            "lookup_bytes": n_pix * LOOKUP_BYTES_PER_PIXEL,
            "build_bytes": 2 * B * C * H * W * 4 + B * H * W * (156 + 78 + 39 + 19) * 4,
            "build_flops": 2 * B * H * W * W * C}


def time_step_kernels(device, peak):
    """Isolated CUDA-event timings (L2 flushed) of this repo's other kernels at the bench shape, with their
    algorithmic bytes: the per-step launch mix is 1 build, 32 fused lookups, 32 upsamplings, 34 + 32 stagings, 64 + 64
    gates, 32 one-channel 7x7 convolutions, 32 flow-head tails and (fp16 iteration) 32 concatenations."""
    import nndepth_b200 as nb
    from nndepth_b200 import _lib
    B, H, W, ch, cx = PAIRS_PER_GPU, 48, 156, 128, 256
    px = B * H * W
    ctot = 2 * (ch + cx)
    cl = torch.channels_last
    torch.manual_seed(4)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=device)
    stream = torch.cuda.current_stream(device)

    def timed(fn, reps=12):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(reps):
            torch.cuda._sleep(1000000)
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        return statistics.median(ts)

    lib = _lib.load()
    sp = _lib.stream_ptr(flush)
    S = torch.empty(B, ctot, H, W, device=device).contiguous(memory_format=cl)
    h = torch.randn(B, ch, H, W, device=device).contiguous(memory_format=cl)
    z = torch.empty_like(h)
    zr = torch.randn(B, 2 * ch, H, W, device=device).contiguous(memory_format=cl)
    q = torch.randn(B, ch, H, W, device=device).contiguous(memory_format=cl)
    bias = torch.randn(2 * ch, device=device)
    motion = torch.randn(B, ch, H, W, device=device).contiguous(memory_format=cl)
    flow = torch.randn(B, 1, H, W, device=device)
    mask = torch.randn(B, 576, H, W, device=device).contiguous(memory_format=cl)
    mbias = torch.randn(576, device=device)
    rows = []

    def add(name, per_step, nbytes, fn):
        us = timed(fn)
        rows.append({"kernel": name, "launches_per_step": per_step, "us_per_launch_l2_flushed": us,
                     "algorithmic_bytes_per_launch": nbytes, "achieved_gbs": nbytes / us / 1e3, "frac_of_hbm_peak": nbytes / us / 1e3 / peak})

    add("gru_gate_r_kernel", 64, px * (2 * ch + ch + ch + 2 * ch) * 4,
        lambda: _lib.check(lib.nnd_gru_gate_r(_lib.ptr(zr), _lib.ptr(bias), _lib.ptr(h), px, ch, _lib.ptr(z), _lib.ptr(S), ctot, sp), "gate_r"))
    add("gru_gate_h_kernel", 64, px * (ch + ch + ch + ch + 2 * ch) * 4,
        lambda: _lib.check(lib.nnd_gru_gate_h(_lib.ptr(q), _lib.ptr(bias), _lib.ptr(z), px, ch, _lib.ptr(h), _lib.ptr(S), ctot, sp), "gate_h"))
    add("gru_stage_cl_kernel (motion features)", 32, px * (ch + 2 * ch) * 4,
        lambda: _lib.check(lib.nnd_gru_stage(_lib.ptr(motion), 1, B, ch, H * W, _lib.ptr(S), ctot, ch + ch, sp), "stage"))
    add("convex_upsample_nhwc8_kernel", 32, px * 576 * 4 + px * 64 * 4 + px * 4,
        lambda: nb.convex_upsample(flow, mask, 8, 0.25, mbias))
    # ---- the fp16 iteration (dense_precision mixed16, the default): same kernels on fp16 staging / pre-activations ----
    from nndepth_b200.raft_stereo import flow_conv7x7_relu, flow_head_tail, nhwc_cat_f16
    S16 = torch.empty(B, ctot, H, W, device=device, dtype=torch.float16).contiguous(memory_format=cl)
    zr16, q16, motion16, mask16 = zr.half(), q.half(), motion.half(), mask.half()
    h16 = torch.empty(B, ch, H, W, device=device, dtype=torch.float16).contiguous(memory_format=cl)
    add("gru_gate_r_f16_kernel", 64, px * (2 * ch * 2 + ch * 4 + ch * 4 + 2 * ch * 2),
        lambda: _lib.check(lib.nnd_gru_gate_r_f16(_lib.ptr(zr16), _lib.ptr(bias), _lib.ptr(h), px, ch, _lib.ptr(z), _lib.ptr(S16), ctot, sp), "gate_r16"))
    add("gru_gate_h_f16_kernel", 64, px * (ch * 2 + ch * 4 + ch * 4 + ch * 4 + 2 * ch * 2 + ch * 2),
        lambda: _lib.check(lib.nnd_gru_gate_h_f16(_lib.ptr(q16), _lib.ptr(bias), _lib.ptr(z), px, ch, _lib.ptr(h), _lib.ptr(S16), ctot,
                                                  _lib.ptr(h16), sp), "gate_h16"))
    add("gru_stage_cl_f16_kernel (fp16 motion features)", 32, px * (ch * 2 + 2 * ch * 2),
        lambda: _lib.check(lib.nnd_gru_stage_f16(_lib.ptr(motion16), 2, B, ch, H * W, _lib.ptr(S16), ctot, ch + ch, sp), "stage16"))
    add("convex_upsample_nhwc8_kernel<fp16 mask>", 32, px * 576 * 2 + px * 64 * 4 + px * 4,
        lambda: nb.convex_upsample(flow, mask16, 8, 0.25, mbias))
    conv7 = torch.nn.Conv2d(1, 128, 7, padding=3).to(device)
    head2 = torch.nn.Conv2d(128, 1, 3, padding=1).to(device)
    coords = torch.randn(B, 1, H, W, device=device)
    cor = torch.randn(B, 192, H, W, device=device).contiguous(memory_format=cl)
    flo16 = torch.randn(B, 64, H, W, device=device).half().contiguous(memory_format=cl)
    with torch.no_grad():
        add("flow_conv7x7_relu_kernel<fp16 out>", 32, px * 4 + px * 128 * 2, lambda: flow_conv7x7_relu(conv7, flow, half=True))
        add("flow_head_tail_kernel<4, fp16 in>", 32, px * 128 * 2 + px * 16, lambda: flow_head_tail(head2, h16, coords, coords))
        add("nhwc_cat_f16_kernel", 32, px * (192 * 4 + 64 * 2 + 256 * 2), lambda: nhwc_cat_f16(cor, flo16))
    return rows


def time_fused_lookup(device, B, out_elem, peak, reps=10):
    """The per-iteration kernel (lookup + convc1 + ReLU) stand-alone at batch ``B`` of the KITTI feature shape, L2 flushed:
    on the row layout with white-noise coordinates (SURVEY 8(d)'s synthetic lookups) and with a smooth disparity field
    like the one the model produces (tools/coords_smoothness.py: <= 4.6 px of disparity range inside any 32-pixel group
    over all 32 iterations), and on the SKEWED copy of the pyramid with the smooth field."""
    import nndepth_b200 as nb
    C, H, W = 256, 48, 156
    torch.manual_seed(3)
    f1, f2 = torch.randn(B, C, H, W, device=device), torch.randn(B, C, H, W, device=device)
    blk = nb.CorrBlock1D(f1, f2, 4, 4)
    del f1, f2
    grid = torch.arange(W, device=device).float().view(1, 1, 1, W).repeat(B, 1, H, 1)
    noise = grid - torch.rand(B, 1, H, W, device=device) * 40
    smooth = grid - torch.nn.functional.interpolate(torch.rand(B, 1, 6, 20, device=device) * 6, size=(H, W), mode="bilinear",
                                                    align_corners=True)
    conv = torch.nn.Conv2d(36, 256, 1).to(device)
    wt = blk.prepare_conv1x1_weight(conv.weight.detach())
    bias = conv.bias.detach()
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=device)
    stream = torch.cuda.current_stream(device)
    nbytes = B * H * W * (164 + 256 * out_elem)

    def timed(coords, skewed, elem=out_elem):
        nbytes = B * H * W * (164 + 256 * elem)

        def launch():
            return blk.lookup_conv1x1(coords, None, bias, relu=True, weight_t=wt, precision="tf32", channels_last=True,
                                      half=elem == 2, skewed=skewed)
        for _ in range(3):
            launch()
        ts = []
        for _ in range(reps):
            flush.fill_(1.0)
            torch.cuda._sleep(200000)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            launch()
            e1.record(stream)
            e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        us = statistics.median(ts)
        res = {"us_per_launch_l2_flushed": us, "achieved_gbs": nbytes / us / 1e3, "frac": nbytes / us / 1e3 / peak}
        if elem != out_elem:
            res["algorithmic_bytes_per_launch"] = nbytes
        return res

    def timed_plain(coords, skewed=False):
        # the drop-in CorrBlock1D.__call__ itself ((B, 36, H, W) fp32 out, 308 B / pixel)
        nb_plain = B * H * W * LOOKUP_BYTES_PER_PIXEL
        for _ in range(3):
            blk(coords, skewed=skewed)
        ts = []
        for _ in range(reps):
            flush.fill_(1.0)
            torch.cuda._sleep(200000)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            blk(coords, skewed=skewed)
            e1.record(stream)
            e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        us = statistics.median(ts)
        return {"us_per_launch_l2_flushed": us, "algorithmic_bytes_per_launch": nb_plain, "achieved_gbs": nb_plain / us / 1e3,
                "frac": nb_plain / us / 1e3 / peak}

    rows_noise = timed(noise, False)
    out = {"batch": B, "pixels": B * H * W, "algorithmic_bytes_per_launch": nbytes, **rows_noise,
           "coords": "white noise: x - U(0, 40), row layout",
           "smooth_field_row_layout": timed(smooth, False), "smooth_field_skewed_layout": timed(smooth, True),
           # the reference's own output dtype (the fp16 output above belongs to the mixed16 step)
           "fp32_output": {"white_noise_row_layout": timed(noise, False, 4), "smooth_field_skewed_layout": timed(smooth, True, 4)},
           "plain_lookup": {"kernels": "corr1d_lookup_lean_kernel<9> (row layout) / corr1d_lookup_skewed_kernel<9>",
                            "white_noise_row_layout": timed_plain(noise), "smooth_field_row_layout": timed_plain(smooth),
                            "smooth_field_skewed_layout": timed_plain(smooth, True),
                            "white_noise_skewed_layout": timed_plain(noise, True)},
           "note": "skewed layout = nnd_corr1d_skew + nnd_corr1d_lookup_conv1x1_skewed (S[j][w1], j = ((w1 >> l) - w2) mod W2_l): "
                   "ncu DRAM read 105 MB vs 268 MB on the row layout for 79 MB of window data (profiles/r2_lookup_ws_b64_summary.txt); "
                   "bit-identical results; pays for smooth fields only"}
    del blk, flush
    torch.cuda.empty_cache()
    return out


def hotpath_configs(skip_cpu=False):
    """The other BASELINE configs (0, 2, 3, 4) through bench_hotpath.py, condensed: kernel microseconds with their
    fraction of the HBM roofline, the reference's ATen chain on torch.cuda beside them, the reference on the host cores."""
    import bench_hotpath as hp
    out = {}
    for name, fn in (("cfg1", hp.cfg1), ("cfg3", hp.cfg3), ("cfg4", hp.cfg4), ("cfg5", hp.cfg5)):
        try:
            line = fn(skip_cpu)
        except Exception as e:          # a failed side leg must not lose the headline line
            out[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
            torch.cuda.empty_cache()
            continue
        roofs = [line.get("roofline")] + [line.get("lookup_roofline")] + list(line.get("other_rooflines", []))
        out[name] = {"workload": line["config"]["workload"], "gpu_us": line["gpu_us"],
                     "rooflines": [{"kernel": r["kernel"], "us": r["us_per_launch"], "algorithmic_bytes": r["algorithmic_bytes_per_launch"],
                                    "achieved_gbs": r["achieved"], "frac": r["frac"]} for r in roofs if r],
                     "gpu_baseline": line.get("gpu_baseline"), "cpu_baseline": line.get("cpu_baseline")}
        torch.cuda.empty_cache()
    return out


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(fused=False):
    """Per-launch DRAM bytes of the per-iteration kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "lookup_traffic.json")) as f:
            return json.load(f).get("fused_dram_bytes_per_launch" if fused else "dram_bytes_per_launch")
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
# main arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    import nndepth_b200 as nb
    from nndepth_b200 import _lib
    from nndepth_b200.engine import OverlappedGather, StereoEngine
    from nndepth_b200.raft_stereo import BaseRAFTStereo

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: nndepth_b200 has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    nb.load_library()
    if args.volume_precision:
        nb.set_volume_precision(args.volume_precision)
    # Parity first: the reference computes in fp32 and the bar is 0.01 px of final end-point error.  The default
    # "mixed16" runs every dense layer on fp16 tensor-core products with fp32 accumulation and gives the convolutions
    # whose rounding would repeat identically in all 32 iterations (ConvGRU, motion encoder, flow head) and the feature
    # encoder two-term weights w_hi + w_lo; tests/test_gpu_dropin.py gates it on weight seeds 0-3 x two inputs, and
    # the EPE of THIS run against the reference model on torch.cuda is measured below (key "parity").  `value` is
    # never measured in the out-of-tolerance "tf32" mode unless asked for explicitly.

    torch.manual_seed(0)
    model = BaseRAFTStereo(iters=ITERS).eval()
    model.dense_precision = args.dense_precision
    engine = StereoEngine(model, device=device, use_cuda_graph=not args.no_graph)
    gen = torch.Generator().manual_seed(1 + rank)
    host_l = (torch.rand((PAIRS_PER_GPU, 3) + IMAGE_HW, generator=gen) * 2 - 1).pin_memory()
    host_r = (torch.rand((PAIRS_PER_GPU, 3) + IMAGE_HW, generator=gen) * 2 - 1).pin_memory()
    dev_l, dev_r = host_l.to(device), host_r.to(device)
    stream = torch.cuda.current_stream(device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    # the only collective: the all-gather of the ranks' disparity maps.  It runs on its own stream behind a snapshot of
    # the maps, overlapped with the next step's forward; the last one is waited for inside the timed region.
    gather = OverlappedGather(world, device=device)

    def step_device():
        disp = engine.infer_device(dev_l, dev_r)
        gather.submit(disp)
        return disp

    pending = []

    def step_host():
        # host buffers in, host result out, two batches in flight: every step issues its own H2D (89 MB) and D2H
        # (15 MB) inside the timed region; they overlap the neighbouring steps' forwards on the copy stream
        pending.append(engine.submit(host_l, host_r))
        if len(pending) > 1:
            return engine.collect(pending.pop(0))
        return None

    def drain_host():
        while pending:
            engine.collect(pending.pop(0))

    def finish_device():
        gather.result()

    def timed(fn, steps, warmup, finish=None):
        for _ in range(warmup):
            fn()
        if finish is not None:
            finish()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            fn()
        if finish is not None:
            finish()            # results of the last in-flight batches are read back inside the timed region
        e1.record(stream)
        torch.cuda.synchronize(device)
        wall = time.perf_counter() - t0
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1), wall * 1e3], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms[0].item(), ms[1].item()

    # launches of this repo's kernels per forward (counted on an eager forward, same path the graph captured);
    # the same eager forward brackets every lookup with CUDA events on the launching stream: the lookup's
    # duration INSIDE the step (pyramid evicted from L2 by the update block's activations between iterations)
    lookup_events = []

    class TimedCorr(nb.CorrBlock1D):
        def _timed(self, fn, coords, *args, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(torch.cuda.current_stream(coords.device))
            out = fn(coords, *args, **kw)
            e1.record(torch.cuda.current_stream(coords.device))
            lookup_events.append((e0, e1))
            return out

        def __call__(self, coords):
            return self._timed(super().__call__, coords)

        def lookup_conv1x1(self, coords, *args, **kw):
            return self._timed(super().lookup_conv1x1, coords, *args, **kw)

    before = _lib.launch_count()
    engine.use_cuda_graph, keep = False, engine.use_cuda_graph
    step_device()
    launches_per_step = _lib.launch_count() - before
    engine.model.corr_fn = TimedCorr
    # queue ~60 ms of GPU spin first: the host then runs ahead of the device for the whole eager step, so the event
    # pairs bracket the kernels back to back on the stream instead of the host's launch gaps
    torch.cuda._sleep(int(1.2e8))
    step_device()
    torch.cuda.synchronize(device)
    in_step_us = sorted(e0.elapsed_time(e1) * 1e3 for e0, e1 in lookup_events)
    engine.model.corr_fn = nb.CorrBlock1D
    engine.use_cuda_graph = keep

    sampler = ClockSampler(physical_gpu_index(local_rank))
    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_device()
    barrier()
    sampler.start()
    dev_ms, _ = timed(step_device, args.steps, 0, finish=finish_device)
    clocks = sampler.stop()
    _, host_wall_ms = timed(step_host, args.steps, 2, finish=drain_host)

    total_pairs = PAIRS_PER_GPU * world * args.steps
    value = total_pairs / (dev_ms / 1e3)
    e2e_value = total_pairs / (host_wall_ms / 1e3)

    # strong scaling (B/n pairs per GPU): the SAME 8 pairs split over the ranks
    strong = None
    if PAIRS_PER_GPU % world == 0:
        per_rank = PAIRS_PER_GPU // world
        if world == 1:
            strong = {"global_pairs": PAIRS_PER_GPU, "pairs_per_gpu": per_rank, "value": value, "unit": UNIT,
                      "ms_per_step": dev_ms / args.steps}
        else:
            sl, sr = dev_l[:per_rank].contiguous(), dev_r[:per_rank].contiguous()
            sgather = OverlappedGather(world, device=device)

            def step_strong():
                sgather.submit(engine.infer_device(sl, sr))

            strong_ms, _ = timed(step_strong, args.steps, 3, finish=sgather.result)
            strong = {"global_pairs": PAIRS_PER_GPU, "pairs_per_gpu": per_rank, "unit": UNIT,
                      "value": PAIRS_PER_GPU * args.steps / (strong_ms / 1e3), "ms_per_step": strong_ms / args.steps}

    # BASELINE configs[4] as it is meant to run: ONE 1080x1920 pair, its epipolar rows split into row bands across the
    # ranks (no halo, no data-path collective; build + 32 lookups per band as one CUDA graph, one all-gather)
    row_band = None
    if not args.skip_hotpath:
        import bench_hotpath as hp
        try:
            row_band = hp.cfg5_sharded(standalone=False)
        except BaseException as e:      # a failed side leg must not lose the headline line
            row_band = {"error": f"{type(e).__name__}: {e}"[:300]} if rank == 0 else None

    line = None
    if rank == 0:
        peak, peak_src = measured_peak()
        kern = time_lookup_kernel(device)
        in_step_med = in_step_us[len(in_step_us) // 2]
        fused_front = bool(getattr(engine.model, "fuse_motion_front", False))
        # fused kernel: windows + coords in, 256 channels out (the 36-channel lookup tensor stays on chip); the output
        # is fp16 when the rest of the iteration runs as fp16 convolutions (mixed16), fp32 otherwise
        out_elem = 2 if args.dense_precision == "mixed16" else 4
        step_bytes = kern["pixels"] * ((4 * 10 * 4 + 4) + 256 * out_elem) if fused_front else kern["lookup_bytes"]
        achieved = step_bytes / (in_step_med * 1e-6) / 1e9
        cpu = None
        if world == 1 and not args.skip_cpu_baseline:
            sec, threads, kind = cpu_forward_seconds(steps=2, warmup=1, pairs=1)
            cpu = {"value": 1.0 / sec, "unit": UNIT, "cores": threads, "kind": kind,
                   "sample": f"1 pair of {IMAGE_HW[0]}x{IMAGE_HW[1]}, {ITERS} iterations, 2 timed forwards, "
                             + ("the reference's BaseRAFTStereo as is (oracle/_ref)" if kind == "reference"
                                else "model shell in torch CPU ops + oracle/torch_port.py correlation")}
        # ---- parity, measured in this run: the timed configuration against the reference model on torch.cuda (strict
        # fp32, same seed-0 weights, this rank's 8 pairs); the same reference run is the like-for-like GPU baseline ----
        parity, gpu_baseline = None, None
        if not args.skip_parity:
            want, ref_ms, ref_kind = reference_on_cuda(device, dev_l, dev_r)
            got = engine.infer_device(dev_l, dev_r)
            per_pair = (got - want).abs().flatten(1).mean(1)
            parity = {"final_epe_px_vs_reference": (got - want).abs().mean().item(), "worst_pair_epe_px": per_pair.max().item(),
                      "bar_px": 0.01, "measured_in_this_run": True, "pairs": int(got.shape[0]),
                      "reference": ("the unmodified reference BaseRAFTStereo (oracle/_ref) on torch.cuda, strict fp32, same weights "
                                    "and inputs" if ref_kind == "reference" else "oracle port on torch.cuda (oracle/_ref missing)"),
                      "mean_abs_disparity_px": want.abs().mean().item(),
                      "other_weight_seeds": "tests/test_gpu_dropin.py::test_headline_mode_against_reference_across_weight_seeds "
                                            "gates seeds 0-3 x (noise, shipped KITTI pair) x batch 8 on the same bar"}
            gpu_baseline = {"what": "the reference model itself (its ATen correlation chain, cuDNN convolutions) on torch.cuda, "
                                    "one B200, batch 8, eager as the reference runs it",
                            "kind": ref_kind, "ms_per_step_strict_fp32": ref_ms["strict_fp32"],
                            "ms_per_step_stock_flags": ref_ms["stock_flags"], "unit": UNIT,
                            "value_strict_fp32": PAIRS_PER_GPU / (ref_ms["strict_fp32"] / 1e3),
                            "value_stock_flags": PAIRS_PER_GPU / (ref_ms["stock_flags"] / 1e3)}
            del want, got
        # ---- the precision-matched number: every dense layer in fp32 like the reference's eval (correlation volume fp32) ----
        value_fp32 = None
        if world == 1 and not args.skip_fp32:
            torch.manual_seed(0)
            m32 = BaseRAFTStereo(iters=ITERS).eval()
            m32.dense_precision = "fp32"
            old_prec = nb.get_volume_precision()
            nb.set_volume_precision("fp32")
            try:
                e32 = StereoEngine(m32, device=device, use_cuda_graph=not args.no_graph)
                for _ in range(2):
                    e32.infer_device(dev_l, dev_r)
                torch.cuda.synchronize(device)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(3):
                    e32.infer_device(dev_l, dev_r)
                e1.record(stream)
                torch.cuda.synchronize(device)
                value_fp32 = {"value": PAIRS_PER_GPU * 3 / (e0.elapsed_time(e1) / 1e3), "unit": UNIT,
                              "ms_per_step": e0.elapsed_time(e1) / 3, "steps": 3,
                              "dtype": "f32 everywhere: cuDNN fp32 convolutions, fp32 FFMA correlation volume"}
            finally:
                nb.set_volume_precision(old_prec)
            del e32, m32
            torch.cuda.empty_cache()
        # ---- option: upsample only the last iteration (what evaluate.py:155 consumes) -- the mask head and the
        # upsampling of the 31 intermediate predictions are skipped; the FINAL disparity is bit-identical (checked here) ----
        value_final_only = None
        if world == 1 and not args.skip_fp32:
            full = engine.infer_device(dev_l, dev_r).clone()
            engine.model.final_only = True
            try:
                for _ in range(3):
                    last = engine.infer_device(dev_l, dev_r)
                torch.cuda.synchronize(device)
                same = bool(torch.equal(last, full))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(args.steps):
                    engine.infer_device(dev_l, dev_r)
                e1.record(stream)
                torch.cuda.synchronize(device)
                value_final_only = {"value": PAIRS_PER_GPU * args.steps / (e0.elapsed_time(e1) / 1e3), "unit": UNIT,
                                    "ms_per_step": e0.elapsed_time(e1) / args.steps, "steps": args.steps,
                                    "final_disparity_bit_identical_to_headline": same,
                                    "note": "NOT the headline: the reference forward returns all 32 upsampled predictions "
                                            "(model.py:130-141) and the headline computes them all"}
            finally:
                engine.model.final_only = False
            del full
        fnet_half = bool(args.dense_precision == "mixed16" and getattr(engine.model, "fp16_encoder", True))
        dense = {"fp32": "cuDNN fp32", "tf32": "cuDNN TF32",
                 "mixed": "ConvGRU cuDNN fp32, other convolutions cuDNN TF32",
                 "mixed2x": "ConvGRU TF32 activations x split fp32 weights [w_hi; w_lo] on tensor cores, other convolutions cuDNN TF32",
                 "mixed16": "fp16 tensor-core products with fp32 accumulation (TF32's 10-bit operand mantissa): ConvGRU fp16 "
                            "activations x two-term weights [w_hi16; w_lo16], fp32 gates and state; motion encoder / flow head "
                            + ("with two-term fp16 weights (w_hi + w_lo); " if getattr(engine.model, "exact_weights", False)
                               else "fp16; ")
                            + "mask head fp16; feature encoder "
                            + (("fp16" + (" with two-term weights" if getattr(engine.model, "exact_encoder", False) else ""))
                               if fnet_half else "cuDNN TF32")
                            + "; one-channel flow convolutions fp32"}[args.dense_precision]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f32 (correlation volume: %s, fp32 accumulate; dense layers: %s)" % (
                ("fp16 x fp16 products of the fp16 encoder's feature maps (exact; what TF32 keeps of the same values)"
                 if (fnet_half and nb.get_volume_precision() == "tf32")
                 else "%s operands rounded to nearest" % nb.get_volume_precision()), dense),
            "dense_precision": args.dense_precision,
            "parity": parity,
            "data": "synthetic", "config": workload_config(world), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": 2 * host_l.numel() * 4, "d2h_bytes_per_step": PAIRS_PER_GPU * IMAGE_HW[0] * IMAGE_HW[1] * 4,
                    "ms_per_step": host_wall_ms / args.steps},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": {"kernel": ("corr1d_lookup_conv1x1_ws_kernel (nnd_corr1d_lookup_conv1x1: lookup + convc1 + ReLU, "
                                    "warp-specialised tcgen05)" if fused_front else "corr1d_lookup_lean_kernel<9> (nnd_corr1d_lookup)")
                                   + ", 32 launches per step",
                         "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(fused_front), "peak_source": peak_src,
                         "how": "CUDA events around each of the 32 per-iteration lookups of one eager step that runs behind a "
                                "queued GPU spin, so the host's launch gaps stay out of the brackets (median); "
                                "algorithmic bytes = (164 B window+coords in + 256 x %d B out) x 59904 pixels for the fused "
                                "kernel, 308 B/pixel for the stand-alone lookup" % (out_elem if fused_front else 4),
                         "us_per_launch_in_step": in_step_med,
                         "algorithmic_bytes_per_launch": step_bytes,
                         "standalone_lookup": {"us_per_launch_l2_flushed": kern["lookup_ms_l2_flushed"] * 1e3,
                                               "us_per_launch_l2_warm": kern["lookup_ms_l2_warm"] * 1e3,
                                               "algorithmic_bytes_per_launch": kern["lookup_bytes"]},
                         "large_batch": time_fused_lookup(device, 64, out_elem, peak) if fused_front else None,
                         "note": "the KITTI batch is a latency-sized launch (405 pixels per SM); large_batch times the same "
                                 "kernel at batch 64, where it is bound by DRAM traffic"},
            "build": {"kernel": "nnd_corr1d_build (%s)" % nb.get_volume_precision(),
                      "us_per_launch_l2_flushed": kern["build_ms_l2_flushed"] * 1e3,
                      "hbm_gbs": kern["build_bytes"] / (kern["build_ms_l2_flushed"] * 1e-3) / 1e9,
                      "frac_of_hbm_peak": kern["build_bytes"] / (kern["build_ms_l2_flushed"] * 1e-3) / 1e9 / peak,
                      "tflops": kern["build_flops"] / (kern["build_ms_l2_flushed"] * 1e-3) / 1e12,
                      "input": "fp32 NCHW feature maps (the reference's), TF32 operands",
                      "fp16_channels_last_input": {
                          "what": "the build the mixed16 step runs: fp16 channels-last maps from the fp16 encoder read in "
                                  "place (K-major fp16 operands, fp32 accumulation); no fp32 / NCHW copies",
                          "us_per_launch_l2_flushed": kern["build_f16_nhwc_ms_l2_flushed"] * 1e3,
                          "algorithmic_bytes_per_launch": kern["build_f16_nhwc_bytes"],
                          "hbm_gbs": kern["build_f16_nhwc_bytes"] / (kern["build_f16_nhwc_ms_l2_flushed"] * 1e-3) / 1e9,
                          "frac_of_hbm_peak": kern["build_f16_nhwc_bytes"] / (kern["build_f16_nhwc_ms_l2_flushed"] * 1e-3) / 1e9 / peak,
                          "tflops": kern["build_flops"] / (kern["build_f16_nhwc_ms_l2_flushed"] * 1e-3) / 1e12}},
            "other_kernels": time_step_kernels(device, peak),
            "cuda_graph": engine.use_cuda_graph,
            "strong": strong,
            "row_band": row_band,
        }
        if value_fp32 is not None:
            line["value_fp32"] = value_fp32
        if value_final_only is not None:
            line["value_final_only"] = value_final_only
        if gpu_baseline is not None:
            line["gpu_baseline"] = gpu_baseline
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if world == 1 and not args.skip_hotpath:
            line["hotpath"] = hotpath_configs(skip_cpu=args.skip_cpu_baseline)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--volume-precision", default=None, choices=["fp32", "tf32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--dense-precision", default="mixed16", choices=["fp32", "mixed", "mixed2x", "mixed16", "tf32"],
                    help="dense layers: 'mixed16' (default) fp16 tensor-core products, fp32 accumulation, two-term weights; "
                         "'mixed2x' ConvGRU with split fp32 weights on TF32 tensor cores + TF32 elsewhere; 'mixed' ConvGRU fp32 + "
                         "TF32 elsewhere; 'fp32' everything fp32; 'tf32' everything TF32 (outside the 0.01 px bar).  The final "
                         "EPE against the reference is measured in the run (key 'parity')")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-parity", action="store_true", help="do not run the reference model on torch.cuda (parity + gpu_baseline)")
    ap.add_argument("--skip-fp32", action="store_true", help="do not time the all-fp32 configuration (value_fp32)")
    ap.add_argument("--skip-hotpath", action="store_true", help="do not time BASELINE configs 0 / 2 / 3 / 4 (hotpath)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

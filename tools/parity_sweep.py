#!/usr/bin/env python
"""Final-disparity EPE of every dense-precision mode of StereoEngine against the UNMODIFIED reference model
(oracle/_ref, torch.cuda, strict fp32), over weight seeds and inputs.  Writes gpurun_out/parity_sweep.json.

    python tools/parity_sweep.py [--seeds 0 1 2 3] [--modes mixed16 mixed2x mixed fp32]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, nargs="+", default=[0, 1, 2, 3])
    ap.add_argument("--modes", nargs="+", default=["mixed16", "mixed2x", "mixed", "fp32"])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_sweep.json"))
    args = ap.parse_args()
    from oracle import ref_shim
    ref_shim.install()
    from nndepth.models.raft_stereo.model import BaseRAFTStereo as RefModel
    from nndepth.data.dataloaders.utils import Padder
    from nndepth_b200.engine import StereoEngine
    from nndepth_b200.raft_stereo import BaseRAFTStereo

    B = args.batch
    gen = torch.Generator().manual_seed(1)
    noise = (torch.rand((B, 3, 375, 1242), generator=gen) * 2 - 1, torch.rand((B, 3, 375, 1242), generator=gen) * 2 - 1)
    kl, kr = ref_shim.kitti_sample_pair()
    kitti = (kl.repeat(B, 1, 1, 1), kr.repeat(B, 1, 1, 1))
    inputs = {"noise": noise, "kitti_pair": kitti}
    rows = []
    for seed in args.seeds:
        torch.manual_seed(seed)
        ref = RefModel(iters=32).eval().cuda()
        state = ref.state_dict()
        want = {}
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        for name, (l, r) in inputs.items():
            l, r = l.cuda(), r.cuda()
            padder = Padder(l.shape[-2:], divis_by=32)
            lp, rp = padder.pad(l, r)
            with torch.no_grad():
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                out = padder.unpad(ref(lp, rp)[-1]["up_disp"])
                torch.cuda.synchronize()
                t_ref = time.perf_counter() - t0
            want[name] = (out, t_ref)
        del ref
        for mode in args.modes:
            model = BaseRAFTStereo(iters=32).eval()
            model.load_state_dict(state)
            base, _, overrides = mode.partition(":")       # e.g. mixed16:exact_encoder=1,fp16_encoder=0
            model.dense_precision = base
            for item in filter(None, overrides.split(",")):
                k, v = item.split("=")
                setattr(model, k, int(v) if k == "dither_weights" else bool(int(v)))
            engine = StereoEngine(model, device="cuda", use_cuda_graph=True)
            for name, (l, r) in inputs.items():
                l, r = l.cuda(), r.cuda()
                got = engine.infer_device(l, r).clone()
                engine.infer_device(l, r)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(3):
                    engine.infer_device(l, r)
                torch.cuda.synchronize()
                ms = (time.perf_counter() - t0) / 3 * 1e3
                ref_out, t_ref = want[name]
                per_pair = (got - ref_out).abs().flatten(1).mean(1)
                row = {"seed": seed, "mode": mode, "input": name, "epe_px": (got - ref_out).abs().mean().item(),
                       "worst_pair_epe_px": per_pair.max().item(), "mean_abs_disp_px": ref_out.abs().mean().item(),
                       "reference_cuda_fp32_forward_s": t_ref, "ms_per_forward": ms}
                rows.append(row)
                print(json.dumps(row), flush=True)
            del engine, model
            torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()

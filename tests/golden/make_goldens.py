"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (``/root/reference`` is mounted there, not on the GPU box)::

    python tests/golden/make_goldens.py

Every array in ``*.npz`` is either a seeded input or the output of a reference class / function
imported from ``/root/reference`` through ``ref_shim``.  The fixtures are small on purpose (they are
committed); full-size parity is checked GPU-vs-oracle on the fly in ``tests/``.

Where the reference does not expose an intermediate (the integer window indices of
``linear_sampler``, raft_stereo/utils.py:16-21), the same ATen expression is evaluated with torch on
CPU here and stored -- flagged ``aten_`` in the key name.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()

from nndepth.models.raft_stereo.cost_volume import CorrBlock1D, GroupCorrBlock1D  # noqa: E402
from nndepth.models.raft_stereo.utils import linear_sampler  # noqa: E402
from nndepth.models.cre_stereo.cost_volume import AGCL  # noqa: E402
from nndepth.models.cre_stereo.utils import bilinear_sampler  # noqa: E402
from nndepth.models.igev_stereo.cost_volume import GeometryAwareCostVolume  # noqa: E402
from nndepth.models.igev_stereo.model import IGEVStereoBase  # noqa: E402
import torch.nn.functional as F  # noqa: E402

torch.set_grad_enabled(False)
torch.set_num_threads(1)          # deterministic summation order for the fixtures


def save(name, **arrays):
    out = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = v
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}.npz  {os.path.getsize(path) / 1024:.1f} KiB  keys={len(out)}")


def x_grid(B, H, W):
    return torch.arange(W).float()[None, None, None, :].repeat(B, 1, H, 1)


def coords_cases(B, H, W, gen):
    """The three coordinate regimes of SURVEY.md 8(d): exact integers, sub-pixel, out of range."""
    grid = x_grid(B, H, W)
    sub = grid - torch.rand(B, 1, H, W, generator=gen) * (W / 4)
    oob = grid + (torch.rand(B, 1, H, W, generator=gen) - 0.5) * (3 * W)
    oob.view(-1)[::7] = -5.25
    oob.view(-1)[3::11] = W + 2.5
    oob.view(-1)[5::13] = float(W - 1)
    return {"int": grid, "sub": sub, "oob": oob}


def aten_indices(coords, widths, radius):
    """raft_stereo/utils.py:16-21 + cost_volume.py:44-46 evaluated by ATen (CPU) per level."""
    out = {}
    n = coords.numel()
    for lvl, w2 in enumerate(widths):
        dx = torch.linspace(-radius, radius, 2 * radius + 1).view(1, -1)
        c = dx + coords.reshape(n, 1) / 2 ** lvl
        c = c / (w2 - 1)
        c = torch.clamp(c, 0, 1)
        c = c * (w2 - 1)
        out[lvl] = (c.floor().type(torch.int64), c.ceil().type(torch.int64))
    return out


def corr1d_case(name, B, C, H, W, seed, levels=4, radius=4):
    gen = torch.Generator().manual_seed(seed)
    f1 = torch.randn(B, C, H, W, generator=gen)
    f2 = torch.randn(B, C, H, W, generator=gen)
    blk = CorrBlock1D(f1, f2, levels, radius)
    arrays = {"fmap1": f1, "fmap2": f2, "num_levels": np.int64(levels), "radius": np.int64(radius)}
    widths = []
    for lvl, p in enumerate(blk.corr_pyramid):
        arrays[f"pyr{lvl}"] = p.reshape(p.shape[0], p.shape[-1])
        widths.append(p.shape[-1])
    for cname, coords in coords_cases(B, H, W, gen).items():
        arrays[f"coords_{cname}"] = coords
        arrays[f"out_{cname}"] = blk(coords)
        for lvl, (i0, i1) in aten_indices(coords, widths[:levels], radius).items():
            arrays[f"aten_i0_{cname}_{lvl}"] = i0.to(torch.int32)
            arrays[f"aten_i1_{cname}_{lvl}"] = i1.to(torch.int32)
    save(name, **arrays)


def sampler_kats():
    """Known-answer tests of ``linear_sampler`` + the integer round-trip tables."""
    arrays = {}
    row = torch.arange(10).float()[None]
    xs = torch.tensor([[-3.2, -0.5, 0, 0.25, 8.75, 9, 9.4, 20]])
    arrays["kat_row"] = row
    arrays["kat_x"] = xs
    arrays["kat_out"] = linear_sampler(row, xs)
    arrays["pool_in"] = torch.arange(7).float()[None]
    arrays["pool_out"] = F.avg_pool1d(torch.arange(7).float()[None, None], 2)[0]
    for w2 in (240, 160, 156, 120, 80, 78, 60, 40, 39, 30, 20, 19, 9, 5, 2):
        x = torch.arange(w2).float()[None]
        t = torch.clamp(x / (w2 - 1), 0, 1) * (w2 - 1)
        arrays[f"aten_rt_floor_{w2}"] = t.floor().to(torch.int32)[0]
        arrays[f"aten_rt_ceil_{w2}"] = t.ceil().to(torch.int32)[0]
        ramp = torch.arange(w2).float()[None] * 1.5 - 3
        arrays[f"rt_out_{w2}"] = linear_sampler(ramp, x)[0]
    save("sampler_kats", **arrays)


def group_corr_case():
    gen = torch.Generator().manual_seed(21)
    B, C, H, W, G = 2, 32, 3, 24, 4
    f1 = torch.randn(B, C, H, W, generator=gen)
    f2 = torch.randn(B, C, H, W, generator=gen)
    blk = GroupCorrBlock1D(f1, f2, 4, 4, G)
    arrays = {"fmap1": f1, "fmap2": f2, "num_groups": np.int64(G)}
    for lvl, p in enumerate(blk.corr_pyramid):
        arrays[f"pyr{lvl}"] = p.reshape(p.shape[0], p.shape[-1])
    for cname, coords in coords_cases(B, H, W, gen).items():
        arrays[f"coords_{cname}"] = coords
        arrays[f"out_{cname}"] = blk(coords)
    save("group_corr1d", **arrays)


def toy_regularizer(vol, feats):
    """Stand-in for the out-of-scope 3-D hourglass: any deterministic map (B,G,D,H,W)->(B,G,D,H,W)."""
    return torch.tanh(vol) * 0.5 + torch.roll(vol, 1, dims=2) * 0.25 + feats[0].mean() * 0.0


def igev_case():
    gen = torch.Generator().manual_seed(31)
    B, C, H, W, G = 2, 72, 3, 24, 8
    f1 = torch.randn(B, C, H, W, generator=gen)
    f2 = torch.randn(B, C, H, W, generator=gen)
    feats = [torch.randn(B, 4, H, W, generator=gen)]
    cv = GeometryAwareCostVolume(f1, f2, feats, toy_regularizer, 4, 4, G)
    vol = cv.build_cost_volume(f1, f2)
    arrays = {"fmap1": f1, "fmap2": f2, "num_groups": np.int64(G), "feat_volume": vol,
              "geo_volume": toy_regularizer(vol.clone().permute(0, 1, 4, 2, 3), feats)}
    # fact 4 of SURVEY.md section 0: channels >= G*G never enter the volume
    f1z, f2z = f1.clone(), f2.clone()
    f1z[:, G * G:] = 0
    f2z[:, G * G:] = 0
    arrays["feat_volume_first64_only"] = cv.build_cost_volume(f1z, f2z)
    for lvl in range(5):
        arrays[f"feat_pyr{lvl}"] = cv.feat_corr_cv[lvl].reshape(-1, cv.feat_corr_cv[lvl].shape[-1])
        arrays[f"geo_pyr{lvl}"] = cv.geo_aware_cv[lvl].reshape(-1, cv.geo_aware_cv[lvl].shape[-1])
    for cname, coords in coords_cases(B, H, W, gen).items():
        arrays[f"coords_{cname}"] = coords
        arrays[f"out_{cname}"] = cv(coords)
    # soft-argmin: softmax over D (model.py:145) then regress_disparity (model.py:92-95)
    z = torch.randn(2, 24, 5, 7, generator=gen) * 3
    z[0, :, 0, 0] = 0.0
    z[0, 5, 0, 1] = 80.0            # one-hot-ish, checks max-subtraction
    z[1, :, 4, 6] = -1e4            # large negative constant row
    p = F.softmax(z, dim=1)
    arrays["sa_logits"] = z
    arrays["sa_softmax"] = p
    arrays["sa_disp"] = IGEVStereoBase.regress_disparity(None, p, 24)
    save("igev", **arrays)


def igev_squeeze_case():
    """cv_squeezer + softmax + regress_disparity exactly as IGEVStereoBase.forward runs them
    (igev_stereo/model.py:65, 143-146) on the reference GeometryAwareCostVolume's geo_aware_cv[0]."""
    gen = torch.Generator().manual_seed(37)
    B, C, H, W, G = 2, 64, 5, 24, 8
    f1 = torch.randn(B, C, H, W, generator=gen)
    f2 = torch.randn(B, C, H, W, generator=gen)
    feats = [torch.randn(B, 4, H, W, generator=gen)]
    cv = GeometryAwareCostVolume(f1, f2, feats, toy_regularizer, 4, 4, G)
    torch.manual_seed(38)
    squeezer = torch.nn.Conv3d(G, 1, 3, 1, 1)
    squeezer.weight.data.mul_(4.0)      # sharper softmax than the default init gives
    geo = cv.geo_aware_cv[0].reshape(B, G, H, W, W).permute(0, 1, 4, 2, 3)
    cost = squeezer(geo).squeeze(1)
    dist = F.softmax(cost, dim=1)
    disp = IGEVStereoBase.regress_disparity(None, dist, W)
    save("igev_squeeze", geo_pyr0=cv.geo_aware_cv[0].reshape(-1, W), weight=squeezer.weight, bias=squeezer.bias,
         cost=cost, disp=disp, shape=np.int64([B, G, H, W, W]))


def agcl_case():
    gen = torch.Generator().manual_seed(41)
    N, C, H, W = 2, 32, 6, 10
    f1 = torch.randn(N, C, H, W, generator=gen)
    f2 = torch.randn(N, C, H, W, generator=gen)
    flow = torch.randn(N, 2, H, W, generator=gen) * 2
    flow[0, :, 0, 0] = torch.tensor([-30.0, 4.0])      # far outside: every corner zero
    flow[0, :, 1, 1] = torch.tensor([0.0, 0.0])        # exact integer position
    flow[1, :, 5, 9] = torch.tensor([0.5, 0.5])        # bottom-right corner straddles the border
    offs = (torch.rand(N, 18, H, W, generator=gen) - 0.5) * 2
    agcl = AGCL(f1, f2)
    arrays = {"fmap1": f1, "fmap2": f2, "flow": flow, "extra_offset": offs}
    for small in (False, True):
        tag = "3x3" if small else "1x9"
        arrays[f"offset_{tag}"] = agcl(flow, offs, small_patch=small, iter_mode=False)
        arrays[f"iter_{tag}"] = agcl(flow, None, small_patch=small, iter_mode=True)
    coords = (agcl.coords + flow).permute(0, 2, 3, 1)
    arrays["warped_right"] = bilinear_sampler(f2, coords)

    # attention hook: any callable on (N, HW, C) pairs (cost_volume.py:91-99)
    def att(left, right):
        return left * 0.5 + right.flip(1) * 0.25, right - left * 0.125

    arrays["offset_att_1x9"] = AGCL(f1, f2, att=att)(flow, offs, small_patch=False, iter_mode=False)
    save("agcl", **arrays)


def state_fingerprint(model):
    """Order-sensitive fingerprint of a state dict: (sum |w|, sum w * ramp) in float64."""
    acc_abs, acc_ramp, count = 0.0, 0.0, 0
    for name, t in model.state_dict().items():
        t = t.detach().double().reshape(-1)
        acc_abs += float(t.abs().sum())
        acc_ramp += float((t * torch.linspace(-1, 1, t.numel(), dtype=torch.float64)).sum())
        count += t.numel()
    return np.float64([acc_abs, acc_ramp, count])


def raft_model_cases():
    """Full-model goldens: the reference BaseRAFTStereo (random init, seed 0) on seeded synthetic pairs.

    The shell model of this repo (nndepth_b200/raft_stereo.py) draws the SAME weights from seed 0 --
    checked here against the reference's state dict, bit for bit -- so only inputs' seeds, the weight
    fingerprint and the reference outputs need to be committed.
    """
    from nndepth.models.raft_stereo.model import BaseRAFTStereo as RefModel
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from nndepth_b200.raft_stereo import BaseRAFTStereo as OurModel

    torch.set_num_threads(os.cpu_count() or 1)
    for name, iters, shape in (("raft_small", 6, (1, 3, 96, 160)), ("raft_kitti", 32, (1, 3, 384, 1248))):
        torch.manual_seed(0)
        ref = RefModel(iters=iters).eval()
        torch.manual_seed(0)
        ours = OurModel(iters=iters).eval()
        rs, os_ = ref.state_dict(), ours.state_dict()
        assert list(rs.keys()) == list(os_.keys()), "state-dict keys differ from the reference"
        for k in rs:
            assert torch.equal(rs[k], os_[k]), f"seeded init differs at {k}"
        gen = torch.Generator().manual_seed(1)
        left = torch.rand(shape, generator=gen) * 2 - 1
        right = torch.rand(shape, generator=gen) * 2 - 1
        outs = ref(left, right)
        arrays = {"iters": np.int64(iters), "shape": np.int64(shape), "fingerprint": state_fingerprint(ref),
                  "final_up_disp": outs[-1]["up_disp"]}
        if name == "raft_small":
            arrays["all_up_disp"] = torch.stack([o["up_disp"] for o in outs])
        save(name, **arrays)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "raft":
        raft_model_cases()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "igev_squeeze":
        igev_squeeze_case()
        sys.exit(0)
    sampler_kats()
    corr1d_case("corr1d_small", B=2, C=32, H=3, W=40, seed=11)
    corr1d_case("corr1d_odd", B=1, C=16, H=2, W=39, seed=12)
    corr1d_case("corr1d_kitti_row", B=1, C=64, H=1, W=156, seed=13)
    corr1d_case("corr1d_r3l3", B=1, C=8, H=2, W=32, seed=14, levels=3, radius=3)
    group_corr_case()
    igev_case()
    igev_squeeze_case()
    agcl_case()
    raft_model_cases()

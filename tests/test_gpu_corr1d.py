"""GPU parity: RAFT-Stereo correlation pyramid + lookup through the C ABI vs the oracle / goldens.

Bars (BASELINE.json): window indices bit-exact vs the CPU reference; lookups on an identical
pyramid bit-exact (same IEEE operation order); fp32 volume within 1e-5 relative (of the volume's
scale -- values cross zero); pooled levels bit-exact given level 0.
"""
import numpy as np
import pytest
import torch

from oracle import corr1d as oc

pytestmark = pytest.mark.gpu

CASES = ["corr1d_small", "corr1d_odd", "corr1d_kitti_row", "corr1d_r3l3"]
REGIMES = ["int", "sub", "oob"]
VOLUME_RTOL = 1e-5      # fp32 bar of BASELINE.json


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def shape_of(g):
    B, _, H, W = g["coords_int"].shape
    return B, H, W


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("regime", REGIMES)
def test_lookup_bit_exact_on_reference_pyramid(golden, case, regime):
    import nndepth_b200 as nb
    g = golden(case)
    L, r = int(g["num_levels"]), int(g["radius"])
    B, H, W = shape_of(g)
    blk = nb.CorrBlock1D.from_pyramid([g[f"pyr{l}"] for l in range(L)], B, H, L, r)
    out = blk(dev(g[f"coords_{regime}"])).cpu().numpy()
    assert out.dtype == np.float32 and out.shape == g[f"out_{regime}"].shape
    np.testing.assert_array_equal(out, g[f"out_{regime}"])


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("regime", REGIMES)
def test_window_indices_bit_exact(golden, case, regime):
    import nndepth_b200 as nb
    g = golden(case)
    L, r = int(g["num_levels"]), int(g["radius"])
    widths = [g[f"pyr{l}"].shape[1] for l in range(L)]
    i0, i1 = nb.lookup_indices(widths, dev(g[f"coords_{regime}"]), L, r)
    for l in range(L):
        np.testing.assert_array_equal(i0[l].cpu().numpy(), g[f"aten_i0_{regime}_{l}"])
        np.testing.assert_array_equal(i1[l].cpu().numpy(), g[f"aten_i1_{regime}_{l}"])


@pytest.mark.parametrize("w2", [240, 160, 156, 120, 80, 78, 60, 40, 39, 30, 20, 19, 9, 5, 2])
def test_integer_round_trip_table(golden, w2):
    """x/(w2-1)*(w2-1) does not round-trip integers; floor/ceil must match ATen's."""
    import nndepth_b200 as nb
    g = golden("sampler_kats")
    coords = torch.arange(w2, dtype=torch.float32, device="cuda").view(1, 1, 1, w2)
    i0, i1 = nb.lookup_indices([w2], coords, 1, 0)
    np.testing.assert_array_equal(i0.view(-1).cpu().numpy(), g[f"aten_rt_floor_{w2}"])
    np.testing.assert_array_equal(i1.view(-1).cpu().numpy(), g[f"aten_rt_ceil_{w2}"])
    ramp = torch.arange(w2, dtype=torch.float32, device="cuda")[None] * 1.5 - 3
    out = nb.linear_sampler(ramp, coords.view(1, w2))
    np.testing.assert_array_equal(out[0].cpu().numpy(), g[f"rt_out_{w2}"])


@pytest.mark.parametrize("w2", [156, 78, 39, 19, 160, 240, 2, 1025])
def test_window_indices_exhaustive_binades(w2):
    """Every float32 in a few binades (and a dense random sweep) -> (floor, ceil) of the clamped
    position must equal numpy's IEEE division, bit for bit.  Guards the kernel's 3-instruction exact
    division (lookup.cu:sampler_quotient) against the generic one."""
    import nndepth_b200 as nb
    rng = np.random.default_rng(w2)
    chunks = []
    for lo in (0.5, 1.0, 8.0, 64.0):                      # all 2^23 floats of [lo, 2*lo)
        start = np.float32(lo).view(np.uint32)
        chunks.append((start + np.arange(1 << 23, dtype=np.uint32)).view(np.float32))
    chunks.append(rng.uniform(-8, w2 + 8, 1 << 22).astype(np.float32))
    chunks.append((rng.standard_normal(1 << 20) * 1e-3).astype(np.float32))
    chunks.append(np.float32([0.0, -0.0, 1e-38, 1e-45, -1e-45, 3e-39, np.inf, -np.inf, 1e30, -1e30, w2 - 1, w2 - 1.5]))
    for xs in chunks:
        coords = torch.from_numpy(xs).cuda().view(1, 1, 1, -1)
        i0, i1 = nb.lookup_indices([w2], coords, 1, 0)
        _, o0, o1 = oc.sampler_indices(xs, w2)
        np.testing.assert_array_equal(i0.view(-1).cpu().numpy(), o0)
        np.testing.assert_array_equal(i1.view(-1).cpu().numpy(), o1)


def test_linear_sampler_known_answers(golden):
    import nndepth_b200 as nb
    g = golden("sampler_kats")
    out = nb.linear_sampler(dev(g["kat_row"]), dev(g["kat_x"]))
    np.testing.assert_array_equal(out.cpu().numpy(), g["kat_out"])
    np.testing.assert_array_equal(out[0].cpu().numpy(), np.float32([0, 0, 0, 0.25, 8.75, 9, 9, 9]))


@pytest.mark.parametrize("case", CASES)
def test_build_fp32_vs_reference(golden, case):
    import nndepth_b200 as nb
    g = golden(case)
    L, r = int(g["num_levels"]), int(g["radius"])
    blk = nb.CorrBlock1D(dev(g["fmap1"]), dev(g["fmap2"]), L, r, precision="fp32")
    pyr = blk.corr_pyramid
    assert len(pyr) == L + 1
    scale = np.abs(g["pyr0"]).max()
    for l in range(L + 1):
        got = pyr[l].reshape(pyr[l].shape[0], -1).cpu().numpy()
        assert got.shape == g[f"pyr{l}"].shape
        np.testing.assert_allclose(got, g[f"pyr{l}"], rtol=VOLUME_RTOL, atol=VOLUME_RTOL * scale)
    # pooling is exact arithmetic on our own level 0: bit-exact against avg_pool1d's (a+b)/2
    lvl = pyr[0].reshape(pyr[0].shape[0], -1).cpu().numpy()
    for l in range(1, L + 1):
        lvl = oc.avg_pool_pairs(lvl)
        np.testing.assert_array_equal(pyr[l].reshape(pyr[l].shape[0], -1).cpu().numpy(), lvl)
    vol = nb.CorrBlock1D.corr(dev(g["fmap1"]), dev(g["fmap2"]), precision="fp32")
    B, H, W = shape_of(g)
    assert tuple(vol.shape) == (B, H, W, W)
    np.testing.assert_array_equal(vol.reshape(-1, W).cpu().numpy(), pyr[0].reshape(-1, W).cpu().numpy())


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("regime", REGIMES)
def test_build_and_lookup_end_to_end(golden, case, regime):
    """Own volume + own lookup vs the reference's output: only the fp32 contraction order differs."""
    import nndepth_b200 as nb
    g = golden(case)
    L, r = int(g["num_levels"]), int(g["radius"])
    blk = nb.CorrBlock1D(dev(g["fmap1"]), dev(g["fmap2"]), L, r, precision="fp32")
    out = blk(dev(g[f"coords_{regime}"])).cpu().numpy()
    scale = np.abs(g["pyr0"]).max()
    np.testing.assert_allclose(out, g[f"out_{regime}"], rtol=VOLUME_RTOL, atol=VOLUME_RTOL * scale)


TF32_RTOL = 1e-3       # tf32-operand bar of BASELINE.json (relative to the volume's scale)


@pytest.mark.parametrize("case", ["corr1d_small", "corr1d_kitti_row", "corr1d_r3l3"])
def test_build_tf32_vs_reference(golden, case):
    """tcgen05 path (TF32 operands, fp32 accumulation in TMEM) against the reference pyramid."""
    import nndepth_b200 as nb
    g = golden(case)
    L, r = int(g["num_levels"]), int(g["radius"])
    blk = nb.CorrBlock1D(dev(g["fmap1"]), dev(g["fmap2"]), L, r, precision="tf32")
    pyr = blk.corr_pyramid
    scale = np.abs(g["pyr0"]).max()
    for l in range(L + 1):
        got = pyr[l].reshape(pyr[l].shape[0], -1).cpu().numpy()
        assert got.shape == g[f"pyr{l}"].shape
        np.testing.assert_allclose(got, g[f"pyr{l}"], rtol=0, atol=TF32_RTOL * scale)
    lvl = pyr[0].reshape(pyr[0].shape[0], -1).cpu().numpy()
    for l in range(1, L + 1):                       # pooling stays exact arithmetic on our own level 0
        lvl = oc.avg_pool_pairs(lvl)
        np.testing.assert_array_equal(pyr[l].reshape(pyr[l].shape[0], -1).cpu().numpy(), lvl)
    for regime in REGIMES:
        out = blk(dev(g[f"coords_{regime}"])).cpu().numpy()
        np.testing.assert_allclose(out, g[f"out_{regime}"], rtol=0, atol=TF32_RTOL * scale)


@pytest.mark.parametrize("shape", [(8, 256, 48, 156, 156), (1, 256, 80, 160, 160), (1, 256, 17, 240, 240),
                                   (2, 24, 3, 300, 300), (1, 16, 2, 520, 264), (1, 40, 2, 40, 72), (3, 7, 5, 8, 12)])
def test_build_tf32_shapes(shape):
    """BASELINE configs 2, 1 and 5 (one row band) at full width, plus multi-tile / ragged shapes."""
    import nndepth_b200 as nb
    B, C, H, W1, W2 = shape
    torch.manual_seed(sum(shape))
    f1 = torch.randn(B, C, H, W1, device="cuda")
    f2 = torch.randn(B, C, H, W2, device="cuda")
    L = 4 if (W2 >> 3) >= 2 else 1
    blk = nb.CorrBlock1D(f1, f2, L, 4, precision="tf32")
    pyr = blk.corr_pyramid
    ref = torch.einsum("bchi,bchj->bhij", f1.double(), f2.double()).float() / C ** 0.5
    scale = ref.abs().max().item()
    assert (pyr[0].reshape(B, H, W1, W2) - ref).abs().max().item() <= TF32_RTOL * scale
    fp32 = nb.CorrBlock1D(f1, f2, L, 4, precision="fp32").corr_pyramid
    for l in range(L):
        lo = pyr[l][:, 0]
        half = lo.shape[1] // 2
        assert torch.equal(pyr[l + 1][:, 0], (lo[:, 0:2 * half:2] + lo[:, 1:2 * half:2]) * 0.5)
        assert (pyr[l][:, 0] - fp32[l][:, 0]).abs().max().item() <= TF32_RTOL * scale


def test_build_tf32_refuses_unaligned_width():
    """TMA needs 16-byte row strides: odd widths are refused loudly, never silently re-routed."""
    import nndepth_b200 as nb
    f = torch.randn(1, 16, 2, 39, device="cuda")
    with pytest.raises(nb.NNDepthError, match="multiples of 4"):
        nb.CorrBlock1D(f, f, 2, 4, precision="tf32")


def test_config1_against_oracle():
    """BASELINE config 1: 256-ch 80x160 features, 4 levels, radius 4 -- oracle (numpy) vs CUDA."""
    import nndepth_b200 as nb
    rng = np.random.default_rng(1)
    B, C, H, W = 1, 256, 80, 160
    f1 = rng.standard_normal((B, C, H, W), dtype=np.float32)
    f2 = rng.standard_normal((B, C, H, W), dtype=np.float32)
    grid = np.broadcast_to(np.arange(W, dtype=np.float32), (B, 1, H, W)).copy()
    sub = grid - rng.uniform(0, 40, size=grid.shape).astype(np.float32)
    sub.reshape(-1)[::97] = -7.5
    sub.reshape(-1)[5::101] = W + 3.25
    ora = oc.CorrBlock1D(f1, f2, 4, 4)
    blk = nb.CorrBlock1D(dev(f1), dev(f2), 4, 4, precision="fp32")
    scale = np.abs(ora.corr_pyramid[0]).max()
    for l in range(4):
        got = blk.corr_pyramid[l].reshape(B * H * W, -1).cpu().numpy()
        np.testing.assert_allclose(got, ora.corr_pyramid[l], rtol=VOLUME_RTOL, atol=VOLUME_RTOL * scale)
    for coords in (grid, sub):
        got = blk(dev(coords)).cpu().numpy()
        np.testing.assert_allclose(got, ora(coords), rtol=VOLUME_RTOL, atol=VOLUME_RTOL * scale)
        # identical pyramid -> identical lookup bits
        same = nb.CorrBlock1D.from_pyramid(ora.corr_pyramid, B, H, 4, 4)
        np.testing.assert_array_equal(same(dev(coords)).cpu().numpy(), ora(coords))
        i0, i1 = blk.lookup_indices(dev(coords))
        for l, (o0, o1) in enumerate(oc.lookup_indices([160, 80, 40, 20], coords, 4, 4)):
            np.testing.assert_array_equal(i0[l].cpu().numpy(), o0)
            np.testing.assert_array_equal(i1[l].cpu().numpy(), o1)


def test_config2_full_size_properties():
    """BASELINE config 2 shapes (B8, 256 ch, 48x156): size-independent properties at full size."""
    import nndepth_b200 as nb
    torch.manual_seed(2)
    B, C, H, W = 8, 256, 48, 156
    f1 = torch.randn(B, C, H, W, device="cuda")
    f2 = torch.randn(B, C, H, W, device="cuda")
    blk = nb.CorrBlock1D(f1, f2, 4, 4, precision="fp32")
    pyr = blk.corr_pyramid
    assert [p.shape[-1] for p in pyr] == [156, 78, 39, 19, 9]
    # (1) the volume against an fp64 contraction on the device
    ref = torch.einsum("bchi,bchj->bhij", f1.double(), f2.double()).float() / 16.0
    scale = ref.abs().max().item()
    err = (pyr[0].reshape(B, H, W, W) - ref).abs().max().item()
    assert err <= VOLUME_RTOL * scale, (err, scale)
    # (2) every pooled level is exactly (x[2j] + x[2j+1]) * 0.5 of the level below, odd tails dropped
    for l in range(4):
        lo = pyr[l][:, 0]
        half = lo.shape[1] // 2
        expect = (lo[:, 0:2 * half:2] + lo[:, 1:2 * half:2]) * 0.5
        assert torch.equal(pyr[l + 1][:, 0], expect)
    # (3) linearity in fmap1: corr(a*f1 + g1, f2) == a*corr(f1,f2) + corr(g1,f2) up to fp32 rounding
    g1 = torch.randn_like(f1)
    lhs = nb.CorrBlock1D.corr(2.0 * f1 + g1, f2, precision="fp32")
    rhs = 2.0 * pyr[0].reshape(B, H, W, W) + nb.CorrBlock1D.corr(g1, f2, precision="fp32")
    assert (lhs - rhs).abs().max().item() <= 4 * VOLUME_RTOL * scale
    # ... and on the tensor-core path (operands rounded to TF32) within the 1e-3 bar
    lhs = nb.CorrBlock1D.corr(2.0 * f1 + g1, f2, precision="tf32")
    rhs = 2.0 * nb.CorrBlock1D.corr(f1, f2, precision="tf32") + nb.CorrBlock1D.corr(g1, f2, precision="tf32")
    assert (lhs - rhs).abs().max().item() <= 4 * TF32_RTOL * scale
    # (4) lookup at far out-of-range coordinates returns the clamped border columns of every level
    far = torch.full((B, 1, H, W), 1e6, device="cuda")
    out = blk(far)
    for l in range(4):
        last = pyr[l][:, 0, -1].reshape(B, H, W)
        for k in range(9):
            assert torch.equal(out[:, l * 9 + k], last)
    out = blk(-far)
    for l in range(4):
        first = pyr[l][:, 0, 0].reshape(B, H, W)
        assert torch.equal(out[:, l * 9 + 4], first)
    # (5) lookup vs a straightforward torch restatement on the device (same IEEE op order)
    coords = torch.arange(W, device="cuda").float().view(1, 1, 1, W).repeat(B, 1, H, 1) - torch.rand(B, 1, H, W, device="cuda") * 40
    got = blk(coords)
    chunks = []
    for l in range(4):
        rows = pyr[l][:, 0]
        w2 = rows.shape[1]
        x = torch.linspace(-4, 4, 9, device="cuda").view(1, 9) + coords.reshape(-1, 1) / 2 ** l
        t = torch.clamp(x / (w2 - 1), 0, 1) * (w2 - 1)
        i0, i1 = t.floor().long(), t.ceil().long()
        coef = i1 - t
        chunks.append((coef * rows.gather(1, i0) + (1 - coef) * rows.gather(1, i1)).view(B, H, W, 9))
    ref_out = torch.cat(chunks, -1).permute(0, 3, 1, 2)
    # ATen's CUDA division multiplies by a reciprocal, so its t can be 1 ulp (~1e-5 at t~100) off the
    # IEEE quotient: values are continuous in t with slope <= ~2*scale per pixel -> compare with that slack
    assert (got - ref_out).abs().max().item() <= 4e-5 * scale


def test_group_corr_block_matches_reference(golden):
    import nndepth_b200 as nb
    g = golden("group_corr1d")
    G = int(g["num_groups"])
    B, H, W = shape_of(g)
    blk = nb.GroupCorrBlock1D(dev(g["fmap1"]), dev(g["fmap2"]), 4, 4, G)
    scale = np.abs(g["pyr0"]).max()
    for l in range(5):
        got = blk.corr_pyramid[l].reshape(-1, g[f"pyr{l}"].shape[1]).cpu().numpy()
        np.testing.assert_allclose(got, g[f"pyr{l}"], rtol=VOLUME_RTOL, atol=VOLUME_RTOL * scale)
    same = nb.GroupCorrBlock1D.from_pyramid([g[f"pyr{l}"] for l in range(4)], B, H, 4, 4, G)
    for regime in REGIMES:
        out = same(dev(g[f"coords_{regime}"])).cpu().numpy()
        np.testing.assert_array_equal(out, g[f"out_{regime}"])
        out = blk(dev(g[f"coords_{regime}"])).cpu().numpy()
        np.testing.assert_allclose(out, g[f"out_{regime}"], rtol=VOLUME_RTOL, atol=VOLUME_RTOL * scale)


def test_edge_cases_and_errors():
    import nndepth_b200 as nb
    f = torch.randn(1, 8, 2, 12, device="cuda")
    # level 3 of a 12-wide volume has width 1: linear_sampler would divide by zero -> refused
    blk = nb.CorrBlock1D(f, f, 4, 4)
    with pytest.raises(nb.NNDepthError, match="width"):
        blk(torch.zeros(1, 1, 2, 12, device="cuda"))
    # a 5-level pyramid of width 12 has an empty level
    with pytest.raises(ValueError):
        nb.CorrBlock1D(f, f, 5, 4)
    # mismatched coords
    ok = nb.CorrBlock1D(f, f, 2, 4)
    with pytest.raises(RuntimeError, match="coords"):
        ok(torch.zeros(1, 1, 3, 12, device="cuda"))
    # W1 != W2 is legal (rectangular volume)
    f2 = torch.randn(1, 8, 2, 20, device="cuda")
    rect = nb.CorrBlock1D(f, f2, 2, 2, precision="fp32")
    assert rect.corr_pyramid[0].shape == (24, 1, 20)
    ref = torch.einsum("bchi,bchj->bhij", f, f2) / 8 ** 0.5
    assert torch.allclose(rect.corr_pyramid[0].reshape(1, 2, 12, 20), ref, rtol=1e-5, atol=1e-5)
    assert rect(torch.zeros(1, 1, 2, 12, device="cuda")).shape == (1, 10, 2, 12)
    # non-contiguous and half inputs are densified/cast like the reference's .float()
    nc = torch.randn(1, 2, 12, 8, device="cuda").permute(0, 3, 1, 2)
    a = nb.CorrBlock1D(nc, nc.half().float(), 2, 4)
    assert a.corr_pyramid[0].shape == (24, 1, 12)
    # NaN coordinates read index 0 instead of faulting
    out = ok(torch.full((1, 1, 2, 12), float("nan"), device="cuda"))
    assert out.shape == (1, 18, 2, 12)
    torch.cuda.synchronize()
    # CorrBlock1D and GroupCorrBlock1D are differentiable (tests/test_gpu_backward.py); a tensor that requires grad is
    # refused loudly only where no backward exists (the fused convolution front of the model shell)
    grp = nb.GroupCorrBlock1D(f.clone().requires_grad_(True), f, 2, 4, 2)
    assert grp._graph_buffer is not None and grp(torch.zeros(1, 1, 2, 12, device="cuda")).requires_grad


def test_config5_row_bands_equal_full_volume():
    """BASELINE config 5 (1080x1920 -> features 136x240): 8 row bands, built and looked up independently
    (one per GPU in the sharded run), reproduce the unsharded pyramid and lookup bit for bit."""
    import nndepth_b200 as nb
    from nndepth_b200.engine import shard_range
    torch.manual_seed(5)
    B, C, H, W = 1, 256, 136, 240
    f1 = torch.randn(B, C, H, W, device="cuda")
    f2 = torch.randn(B, C, H, W, device="cuda")
    coords = (torch.arange(W, device="cuda").float().view(1, 1, 1, W).repeat(B, 1, H, 1)
              - torch.rand(B, 1, H, W, device="cuda") * 60)
    full = nb.CorrBlock1D(f1, f2, 4, 4)
    full_out = full(coords)
    full_pyr = [p.reshape(B, H, W, -1) for p in full.corr_pyramid[:4]]
    for rank in range(8):
        h0, h1 = shard_range(H, rank, 8)
        band = nb.CorrBlock1D(f1[:, :, h0:h1].contiguous(), f2[:, :, h0:h1].contiguous(), 4, 4)
        for l, p in enumerate(band.corr_pyramid[:4]):
            assert torch.equal(p.reshape(B, h1 - h0, W, -1), full_pyr[l][:, h0:h1])
        assert torch.equal(band(coords[:, :, h0:h1].contiguous()), full_out[:, :, h0:h1])


@pytest.mark.parametrize("shape,c_out", [((8, 48, 156), 256), ((1, 5, 52), 64), ((2, 3, 40), 7)])
def test_lookup_conv1x1_fusion(shape, c_out):
    """Lookup fused with the motion encoder's 1x1 convolution + ReLU vs the two-step path on the device."""
    import nndepth_b200 as nb
    B, H, W = shape
    torch.manual_seed(B + H + W)
    f1 = torch.randn(B, 32, H, W, device="cuda")
    f2 = torch.randn(B, 32, H, W, device="cuda")
    blk = nb.CorrBlock1D(f1, f2, 4, 4)
    coords = torch.arange(W, device="cuda").float().view(1, 1, 1, W).repeat(B, 1, H, 1) - torch.rand(B, 1, H, W, device="cuda") * 20
    conv = torch.nn.Conv2d(36, c_out, 1).cuda()
    looked = blk(coords)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            ref = torch.relu(torch.einsum("ok,bkhw->bohw", conv.weight.view(c_out, 36).double(), looked.double())
                             + conv.bias.double().view(1, -1, 1, 1)).float()
            got = blk.lookup_conv1x1(coords, conv.weight, conv.bias, relu=True)
            lin = blk.lookup_conv1x1(coords, conv.weight, None, relu=False)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    scale = ref.abs().max().item()
    assert got.shape == (B, c_out, H, W)
    assert (got - ref).abs().max().item() <= 1e-5 * scale
    ref_lin = torch.einsum("ok,bkhw->bohw", conv.weight.view(c_out, 36).double(), looked.double()).float()
    assert (lin - ref_lin).abs().max().item() <= 1e-5 * scale
    # tensor-core variant: operands rounded to nearest TF32 -> the 1e-3 bar
    with torch.no_grad():
        tc = blk.lookup_conv1x1(coords, conv.weight, conv.bias, relu=True, precision="tf32")
        tc_cl = blk.lookup_conv1x1(coords, conv.weight, conv.bias, relu=True, precision="tf32", channels_last=True)
    assert (tc - ref).abs().max().item() <= 1e-3 * scale
    assert tc_cl.shape == tc.shape and tc_cl.is_contiguous(memory_format=torch.channels_last)
    if c_out == 256:
        # c_out = 256 channels-last runs on the tcgen05 kernel: same TF32 operands, another summation order
        assert (tc_cl - ref).abs().max().item() <= 1e-3 * scale
        assert (tc_cl - tc).abs().max().item() <= 1e-5 * scale
    else:
        assert torch.equal(tc_cl, tc)                   # same kernel, same numbers, channels-last memory
    if c_out <= 256:
        with torch.no_grad():
            tc_h = blk.lookup_conv1x1(coords, conv.weight, conv.bias, relu=True, precision="tf32", channels_last=True, half=True)
        assert tc_h.dtype == torch.float16 and tc_h.permute(0, 2, 3, 1).is_contiguous()
        assert torch.equal(tc_h, tc_cl.half())          # the same values rounded to nearest fp16


@pytest.mark.parametrize("shape", [(8, 48, 156), (1, 5, 52), (3, 7, 44), (1, 1, 16), (20, 48, 156)])
def test_lookup_conv1x1_tcgen05_path(shape):
    """The warp-specialised tcgen05 / TMEM kernel (c_out = 256, channels-last: the shipping path) against the mma.sync
    kernel (NCHW output): same TF32 operands, another summation order.  CTAs own contiguous pixel ranges that straddle
    images, the last tile of a range is ragged, and the largest shape runs 8 tiles per CTA (every mbarrier phase wraps)."""
    import nndepth_b200 as nb
    B, H, W = shape
    torch.manual_seed(B * H + W)
    f1, f2 = torch.randn(B, 64, H, W, device="cuda"), torch.randn(B, 64, H, W, device="cuda")
    blk = nb.CorrBlock1D(f1, f2, 4, 4)
    coords = torch.arange(W, device="cuda").float().view(1, 1, 1, W).repeat(B, 1, H, 1) - torch.rand(B, 1, H, W, device="cuda") * 30
    coords.view(-1)[::17] = -7.5
    coords.view(-1)[5::23] = W + 3.25
    conv = torch.nn.Conv2d(36, 256, 1).cuda()
    with torch.no_grad():
        base = blk.lookup_conv1x1(coords, conv.weight, conv.bias, relu=True, precision="tf32")
        base_lin = blk.lookup_conv1x1(coords, conv.weight, None, relu=False, precision="tf32")
        for _ in range(2):                  # twice: barriers / TMEM are re-initialised per launch
            umma = blk.lookup_conv1x1(coords, conv.weight, conv.bias, relu=True, precision="tf32", channels_last=True)
        umma_h = blk.lookup_conv1x1(coords, conv.weight, conv.bias, relu=True, precision="tf32", channels_last=True, half=True)
        lin = blk.lookup_conv1x1(coords, conv.weight, None, relu=False, precision="tf32", channels_last=True)
    scale = base_lin.abs().max().item()
    assert umma.is_contiguous(memory_format=torch.channels_last) and umma.shape == base.shape
    assert (umma - base).abs().max().item() <= 1e-5 * scale
    assert (lin - base_lin).abs().max().item() <= 1e-5 * scale
    assert torch.equal(umma_h, umma.half())


@pytest.mark.parametrize("shape", [(8, 48, 156, 156), (1, 5, 52, 52), (3, 7, 44, 60), (2, 3, 60, 44), (1, 1, 16, 16), (20, 48, 156, 156)])
@pytest.mark.parametrize("field", ["smooth", "noise"])
def test_lookup_conv1x1_skewed_layout_is_bit_identical(shape, field):
    """The fused kernel reading the SKEWED copy of the pyramid (S[j][w1], j = ((w1 >> l) - w2) mod W2_l) against the same
    kernel on the row layout: identical bits for smooth and for white-noise coordinates, out-of-range coordinates,
    rectangular volumes (W1 != W2), groups that straddle epipolar rows, ragged last tiles."""
    import nndepth_b200 as nb
    B, H, W1, W2 = shape
    torch.manual_seed(B * H + W1 + W2)
    f1, f2 = torch.randn(B, 64, H, W1, device="cuda"), torch.randn(B, 64, H, W2, device="cuda")
    blk = nb.CorrBlock1D(f1, f2, 4, 4)
    grid = torch.arange(W1, device="cuda").float().view(1, 1, 1, W1).repeat(B, 1, H, 1)
    if field == "smooth":
        bump = torch.nn.functional.interpolate(torch.rand(B, 1, 2, 3, device="cuda") * 12, size=(H, W1), mode="bilinear",
                                               align_corners=True)
        coords = grid - bump
    else:
        coords = grid - torch.rand(B, 1, H, W1, device="cuda") * 30
    coords.view(-1)[::17] = -7.5
    coords.view(-1)[5::23] = W2 + 3.25
    conv = torch.nn.Conv2d(36, 256, 1).cuda()
    with torch.no_grad():
        for half in (False, True):
            rows = blk.lookup_conv1x1(coords, conv.weight, conv.bias, relu=True, precision="tf32", channels_last=True, half=half)
            skew = blk.lookup_conv1x1(coords, conv.weight, conv.bias, relu=True, precision="tf32", channels_last=True, half=half,
                                      skewed=True)
            assert skew.dtype == rows.dtype and skew.is_contiguous(memory_format=torch.channels_last)
            assert torch.equal(skew, rows)
        # the plain lookup (CorrBlock1D.__call__) on the skewed copy: the same (B, 36, H, W) tensor bit for bit
        plain, plain_skew = blk(coords), blk(coords, skewed=True)
        assert plain_skew.shape == plain.shape and plain_skew.is_contiguous()
        assert torch.equal(plain_skew, plain)


def test_randomised_lookup_sweep_bit_exact():
    """40 random (shape, levels, radius, coordinate regime) draws: lookup on an identical pyramid and the integer
    window indices must equal the oracle bit for bit -- covers the generic kernel (radius != 4, ragged widths,
    unaligned pitches) as well as the RAFT fast path."""
    import nndepth_b200 as nb
    rng = np.random.default_rng(2026)
    for trial in range(40):
        B, H = int(rng.integers(1, 3)), int(rng.integers(1, 4))
        W1 = int(rng.integers(2, 70))
        L = int(rng.integers(1, 5))
        W2 = int(rng.integers(2 << (L - 1), 90))          # the last used level keeps width >= 2
        r = int(rng.choice([0, 1, 2, 3, 4, 4, 4, 5, 7]))
        vol = rng.standard_normal((B * H * W1, W2)).astype(np.float32)
        pyr = oc.build_pyramid(vol, L)
        regime = trial % 4
        base = np.broadcast_to(np.arange(W1, dtype=np.float32), (B, 1, H, W1))
        if regime == 0:
            coords = base.copy()                                              # exact integers
        elif regime == 1:
            coords = base - rng.uniform(0, W2, size=base.shape).astype(np.float32)
        elif regime == 2:
            coords = rng.uniform(-3 * W2, 4 * W2, size=base.shape).astype(np.float32)   # mostly out of range
        else:
            coords = (rng.integers(0, W2, size=base.shape) + rng.choice([0.0, 0.5, 1e-7, -1e-7], size=base.shape)).astype(np.float32)
        coords = np.ascontiguousarray(coords, dtype=np.float32)
        blk = nb.CorrBlock1D.from_pyramid(pyr, B, H, L, r)
        got = blk(dev(coords)).cpu().numpy()
        want = oc.lookup(pyr, coords, L, r)
        np.testing.assert_array_equal(got, want, err_msg=f"trial {trial}: B{B} H{H} W1={W1} W2={W2} L{L} r{r} regime {regime}")
        i0, i1 = blk.lookup_indices(dev(coords))
        for lvl, (o0, o1) in enumerate(oc.lookup_indices([W2 >> l for l in range(L)], coords, L, r)):
            np.testing.assert_array_equal(i0[lvl].cpu().numpy(), o0)
            np.testing.assert_array_equal(i1[lvl].cpu().numpy(), o1)


@pytest.mark.parametrize("G", [3, 5, 6, 12])
def test_group_block_any_group_count_against_reference(G):
    """The reference accepts any ``num_groups`` with G * G <= C (its split quirk reads G chunks of G channels,
    cost_volume.py:115-121); so does the kernel (group sizes 1..16 are instantiated)."""
    ref_shim = pytest.importorskip("oracle.ref_shim")
    if not ref_shim.available():
        pytest.skip("the vendored reference (oracle/_ref) is not staged")
    ref_shim.install()
    import nndepth_b200 as nb
    from nndepth.models.raft_stereo.cost_volume import GroupCorrBlock1D as RefGroup
    B, C, H, W = 2, 160, 3, 40
    torch.manual_seed(G)
    f1, f2 = torch.randn(B, C, H, W, device="cuda"), torch.randn(B, C, H, W, device="cuda")
    grid = torch.arange(W, device="cuda").float().view(1, 1, 1, W).repeat(B, 1, H, 1)
    coords = grid - torch.rand(B, 1, H, W, device="cuda") * 12
    old = nb.get_volume_precision()
    nb.set_volume_precision("fp32")
    try:
        got = nb.GroupCorrBlock1D(f1, f2, 4, 4, G)(coords)
    finally:
        nb.set_volume_precision(old)
    want = RefGroup(f1.double(), f2.double(), 4, 4, G)(coords.double())
    assert got.shape == want.shape
    scale = want.abs().max().item()
    assert (got.double() - want).abs().max().item() <= 1e-5 * scale


@pytest.mark.parametrize("shape", [(8, 256, 48, 156, 156), (1, 256, 80, 160, 160), (2, 64, 5, 52, 44), (1, 96, 3, 260, 300), (3, 8, 2, 16, 16)])
def test_build_from_fp16_channels_last_maps(shape):
    """fp16 channels-last feature maps (a cuDNN fp16 encoder's output) are contracted where they lie: K-major fp16
    operands on tcgen05.  Against the fp64 contraction of the same fp16 values (1e-5 of the volume's scale: the products
    are exact, only the fp32 accumulation order differs) and against the TF32 build of the values converted to fp32
    NCHW; the pooled levels are bit-exact poolings of level 0."""
    import nndepth_b200 as nb
    B, C, H, W1, W2 = shape
    torch.manual_seed(W1 + C)
    f1 = torch.randn(B, C, H, W1, device="cuda").half().contiguous(memory_format=torch.channels_last)
    f2 = torch.randn(B, C, H, W2, device="cuda").half().contiguous(memory_format=torch.channels_last)
    blk = nb.CorrBlock1D(f1, f2, 4, 4)
    assert blk._pyr.levels[0].shape[0] == B * H * W1
    want = torch.einsum("bchi,bchj->bhij", f1.double(), f2.double()) / C ** 0.5
    got = blk.corr_pyramid[0].reshape(B, H, W1, W2)
    scale = want.abs().max().item()
    assert (got.double() - want).abs().max().item() <= 1e-5 * scale
    ref = nb.CorrBlock1D(f1.float().contiguous(), f2.float().contiguous(), 4, 4)
    for a, b in zip(blk.corr_pyramid[:4], ref.corr_pyramid[:4]):
        assert a.shape == b.shape
        assert (a - b).abs().max().item() <= 1e-5 * scale
    for lo, hi in zip(blk.corr_pyramid[:3], blk.corr_pyramid[1:4]):
        w = hi.shape[-1]
        assert torch.equal(hi, (lo[..., 0:2 * w:2] + lo[..., 1:2 * w:2]) / 2)
    # the lookup reads the same storage
    coords = torch.arange(W1, device="cuda").float().view(1, 1, 1, W1).repeat(B, 1, H, 1) - 3.3
    assert (blk(coords) - ref(coords)).abs().max().item() <= 1e-5 * scale

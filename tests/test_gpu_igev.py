"""GPU parity: IGEV group-wise volume, geometry re-layout + pooling, dual lookup, soft-argmin."""
import numpy as np
import pytest
import torch

from oracle import igev as oi

pytestmark = pytest.mark.gpu

REGIMES = ["int", "sub", "oob"]
VOLUME_RTOL = 1e-5          # fp32 bar of BASELINE.json
SOFTARGMIN_ATOL = 2e-4      # px; far inside the 0.01 px end-point budget of BASELINE.json


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def toy_regularizer(vol, feats):
    """Same stand-in as tests/golden/make_goldens.py (the 3-D hourglass is out of scope)."""
    return torch.tanh(vol) * 0.5 + torch.roll(vol, 1, dims=2) * 0.25 + feats[0].mean() * 0.0


def test_groupwise_volume_matches_reference(golden):
    import nndepth_b200 as nb
    g = golden("igev")
    G = int(g["num_groups"])
    f1, f2 = dev(g["fmap1"]), dev(g["fmap2"])
    cv = nb.GeometryAwareCostVolume(f1, f2, [dev(np.zeros((2, 4, 3, 24), np.float32))], toy_regularizer, 4, 4, G)
    vol = cv.build_cost_volume(f1, f2)
    ref = g["feat_volume"]
    assert tuple(vol.shape) == ref.shape
    np.testing.assert_allclose(vol.cpu().numpy(), ref, rtol=VOLUME_RTOL, atol=VOLUME_RTOL * np.abs(ref).max())
    # only the first G*G channels enter the volume
    f1z = f1.clone()
    f1z[:, G * G:] = 7.0
    assert torch.equal(cv.build_cost_volume(f1z, f2), vol)
    # pyramids of the constructor: feature side within tolerance, pooling exact on own level 0
    scale = np.abs(ref).max()
    for l in range(5):
        got = cv.feat_corr_cv[l].reshape(-1, g[f"feat_pyr{l}"].shape[1]).cpu().numpy()
        np.testing.assert_allclose(got, g[f"feat_pyr{l}"], rtol=VOLUME_RTOL, atol=VOLUME_RTOL * scale)
    # geometry side: regulariser is elementwise-ish on our volume -> tolerance, and pooling exact
    for l in range(5):
        got = cv.geo_aware_cv[l].reshape(-1, g[f"geo_pyr{l}"].shape[1]).cpu().numpy()
        np.testing.assert_allclose(got, g[f"geo_pyr{l}"], rtol=1e-4, atol=1e-5 * scale)
    # model.py:144 reads geo_aware_cv[0] and reshapes it
    B, H, W = 2, 3, 24
    assert cv.geo_aware_cv[0].reshape(B, G, H, W, W).permute(0, 1, 4, 2, 3).shape == (B, G, W, H, W)


def test_geo_transpose_pool_bit_exact(golden):
    """Pure data movement + avg_pool1d arithmetic: bit-exact against the reference pyramid."""
    import nndepth_b200 as nb
    g = golden("igev")
    G = int(g["num_groups"])
    geo = dev(g["geo_volume"])

    def regulariser(vol, feats):
        return geo

    cv = nb.GeometryAwareCostVolume(dev(g["fmap1"]), dev(g["fmap2"]), [], regulariser, 4, 4, G)
    for l in range(5):
        got = cv.geo_aware_cv[l].reshape(-1, g[f"geo_pyr{l}"].shape[1]).cpu().numpy()
        np.testing.assert_array_equal(got, g[f"geo_pyr{l}"])


@pytest.mark.parametrize("regime", REGIMES)
def test_dual_lookup_bit_exact(golden, regime):
    import nndepth_b200 as nb
    g = golden("igev")
    G = int(g["num_groups"])
    cv = nb.GeometryAwareCostVolume.from_pyramids([g[f"feat_pyr{l}"] for l in range(4)],
                                                  [g[f"geo_pyr{l}"] for l in range(4)], 2, 3, 4, 4, G)
    out = cv(dev(g[f"coords_{regime}"])).cpu().numpy()
    assert out.shape == g[f"out_{regime}"].shape == (2, 4 * 2 * G * 9, 3, 24)
    np.testing.assert_array_equal(out, g[f"out_{regime}"])


def test_soft_argmin_golden(golden):
    import nndepth_b200 as nb
    g = golden("igev")
    out = nb.soft_argmin(dev(g["sa_logits"])).cpu().numpy()
    assert out.shape == g["sa_disp"].shape
    np.testing.assert_allclose(out, g["sa_disp"], rtol=1e-5, atol=SOFTARGMIN_ATOL)
    assert out[0, 0, 0, 1] == -5.0                                   # one-hot row
    np.testing.assert_allclose(out[0, 0, 0, 0], -11.5, rtol=1e-6)    # uniform row: mean of 0..23
    np.testing.assert_allclose(out[1, 0, 4, 6], -11.5, rtol=1e-6)    # large negative constant row


@pytest.mark.parametrize("shape", [(2, 160, 30, 40), (1, 7, 3, 5), (3, 33, 1, 70), (1, 1, 2, 2)])
def test_soft_argmin_vs_oracle(shape):
    import nndepth_b200 as nb
    rng = np.random.default_rng(5)
    z = (rng.standard_normal(shape) * 4).astype(np.float32)
    out = nb.soft_argmin(dev(z)).cpu().numpy()
    np.testing.assert_allclose(out, oi.soft_argmin(z), rtol=1e-5, atol=SOFTARGMIN_ATOL)


def test_config4_shapes_properties():
    """BASELINE config 4 geometry at reduced batch (B=2 of 16; rows are independent): 120x160, G=8, D=160."""
    import nndepth_b200 as nb
    torch.manual_seed(4)
    B, C, H, W, G = 2, 256, 120, 160, 8
    f1 = torch.randn(B, C, H, W, device="cuda")
    f2 = torch.randn(B, C, H, W, device="cuda")

    def regulariser(vol, feats):
        return torch.tanh(vol) * 0.5 + 0.125

    cv = nb.GeometryAwareCostVolume(f1, f2, [], regulariser, 4, 4, G)
    feat, geo = cv.feat_corr_cv, cv.geo_aware_cv
    # volume vs fp64 contraction over the first 64 channels
    a = f1[:, :G * G].reshape(B, G, G, H, W).double()
    b = f2[:, :G * G].reshape(B, G, G, H, W).double()
    ref = (torch.einsum("bgchi,bgchj->bghij", a, b) / G ** 0.5).float()
    scale = ref.abs().max().item()
    assert (feat[0].reshape(B, G, H, W, W) - ref).abs().max().item() <= VOLUME_RTOL * scale
    # geometry level 0 == regulariser(vol permuted) permuted back, bit for bit
    expect = regulariser(feat[0].reshape(B, G, H, W, W).permute(0, 1, 4, 2, 3), []).permute(0, 1, 3, 4, 2)
    assert torch.equal(geo[0].reshape(B, G, H, W, W), expect)
    for pyr in (feat, geo):
        assert [p.shape[-1] for p in pyr] == [160, 80, 40, 20, 10]
        for l in range(4):
            lo = pyr[l][:, 0]
            assert torch.equal(pyr[l + 1][:, 0], (lo[:, 0::2] + lo[:, 1::2]) * 0.5)
    # dual lookup vs the single-pyramid kernel applied per (source, group) plane
    coords = torch.arange(W, device="cuda").float().view(1, 1, 1, W).repeat(B, 1, H, 1) - torch.rand(B, 1, H, W, device="cuda") * 40
    out = cv(coords)
    assert out.shape == (B, 576, H, W)
    for src, pyr in enumerate((feat, geo)):
        for g in (0, 3, 7):
            levels = [p[:, 0].reshape(B, G, H * W, -1)[:, g].reshape(B * H * W, -1) for p in pyr[:4]]
            single = nb.CorrBlock1D.from_pyramid(levels, B, H, 4, 4)(coords)
            for l in range(4):
                ch = l * 144 + src * 72 + g * 9
                assert torch.equal(out[:, ch:ch + 9], single[:, l * 9:(l + 1) * 9])
    # soft-argmin at the config's (D,H,W) vs torch on the device
    z = torch.randn(B, W, H, W, device="cuda") * 3
    d = torch.arange(W, device="cuda").float().view(1, -1, 1, 1)
    ref_sa = -(torch.softmax(z.double(), 1) * d).sum(1, keepdim=True).float()
    assert (nb.soft_argmin(z) - ref_sa).abs().max().item() <= SOFTARGMIN_ATOL


@pytest.mark.parametrize("shape,levels", [((1, 64, 3, 20), 2), ((1, 64, 2, 24), 2), ((2, 64, 2, 32), 3), ((1, 64, 2, 12), 1)])
def test_constructor_and_lookup_vs_oracle(shape, levels):
    """Both storage layouts against the oracle: W2 = 20 / 12 are not multiples of 8 -> reference row layout and
    the generic kernels; W2 = 24 / 32 -> interleaved pyramids with fewer than four levels."""
    import nndepth_b200 as nb
    rng = np.random.default_rng(11)
    B, C, H, W = shape
    G = 8
    f1 = rng.standard_normal(shape, dtype=np.float32)
    f2 = rng.standard_normal(shape, dtype=np.float32)

    def reg_np(vol):
        return np.tanh(vol) * 0.5 + 0.25

    cv = nb.GeometryAwareCostVolume(dev(f1), dev(f2), [], lambda vol, feats: torch.tanh(vol) * 0.5 + 0.25, levels, 4, G)
    assert cv._interleaved == (W % 8 == 0)
    vol = oi.groupwise_volume(f1, f2, G)
    feat_pyr, geo_pyr = oi.volume_pyramids(vol, reg_np(vol.transpose(0, 1, 4, 2, 3)), levels)
    scale = np.abs(vol).max()
    for l in range(levels + 1):
        np.testing.assert_allclose(cv.feat_corr_cv[l].reshape(feat_pyr[l].shape).cpu().numpy(), feat_pyr[l],
                                   rtol=VOLUME_RTOL, atol=VOLUME_RTOL * scale)
        np.testing.assert_allclose(cv.geo_aware_cv[l].reshape(geo_pyr[l].shape).cpu().numpy(), geo_pyr[l],
                                   rtol=1e-4, atol=1e-5 * scale)
    coords = (np.broadcast_to(np.arange(W, dtype=np.float32), (B, 1, H, W)) - rng.uniform(-3, 9, (B, 1, H, W))).astype(np.float32)
    same = nb.GeometryAwareCostVolume.from_pyramids(feat_pyr[:levels], geo_pyr[:levels], B, H, levels, 4, G)
    np.testing.assert_array_equal(same(dev(coords)).cpu().numpy(), oi.gev_lookup(feat_pyr, geo_pyr, coords, levels, 4, G))
    got = cv(dev(coords)).cpu().numpy()
    np.testing.assert_allclose(got, oi.gev_lookup(feat_pyr, geo_pyr, coords, levels, 4, G), rtol=1e-4, atol=1e-4 * scale)


def test_squeeze_soft_argmin_golden(golden):
    """Fused cv_squeezer + soft-argmin against the reference model's own Conv3d / softmax / regress chain."""
    import nndepth_b200 as nb
    g = golden("igev_squeeze")
    B, G, H, W1, W2 = (int(v) for v in g["shape"])
    geo0 = g["geo_pyr0"]
    pyr = oi.volume_pyramids(geo0.reshape(B, G, H, W1, W2), geo0.reshape(B, G, H, W1, W2).transpose(0, 1, 4, 2, 3), 4)[1]
    cv = nb.GeometryAwareCostVolume.from_pyramids(pyr[:4], pyr[:4], B, H, 4, 4, G)
    assert cv._interleaved
    sq = torch.nn.Conv3d(G, 1, 3, 1, 1).cuda()
    sq.weight.data.copy_(dev(g["weight"]))
    sq.bias.data.copy_(dev(g["bias"]))
    disp, cost = cv.init_disparity(sq, return_cost=True)
    scale = np.abs(g["cost"]).max()
    np.testing.assert_allclose(cost.cpu().numpy(), g["cost"], rtol=VOLUME_RTOL, atol=VOLUME_RTOL * scale)
    np.testing.assert_allclose(disp.cpu().numpy(), g["disp"], rtol=0, atol=SOFTARGMIN_ATOL)
    np.testing.assert_array_equal(cv.init_disparity(sq).cpu().numpy(), disp.cpu().numpy())


@pytest.mark.parametrize("shape", [(1, 3, 5, 40), (2, 4, 12, 64), (1, 7, 9, 168), (1, 2, 6, 264), (3, 1, 1, 8)])
def test_squeeze_soft_argmin_vs_oracle(shape):
    """Through the C ABI on a hand-interleaved volume: ragged tiles in both directions (H odd, W1 % 4 != 0),
    D not a multiple of 32, the 512-thread variant (D = 264), a single-pixel image, no bias (D = 64)."""
    from nndepth_b200 import _lib
    rng = np.random.default_rng(5)
    B, H, W1, D = shape
    G = 8
    geo = rng.standard_normal((B, G, H, W1, D), dtype=np.float32)
    weight = (rng.standard_normal((1, G, 3, 3, 3)) * 0.3).astype(np.float32)
    bias = None if D == 64 else np.float32([0.37])
    il = dev(geo.transpose(0, 2, 3, 4, 1))                     # [b][h][w1][d][g]
    w_d = dev(weight)
    b_d = dev(bias) if bias is not None else None
    disp = torch.empty(B, 1, H, W1, device="cuda")
    cost = torch.empty(B, D, H, W1, device="cuda")
    _lib.check(_lib.load().nnd_gev_squeeze_soft_argmin(_lib.ptr(il), _lib.ptr(w_d), _lib.ptr(b_d) if bias is not None else None,
                                                       B, G, D, H, W1, _lib.ptr(disp), _lib.ptr(cost), _lib.stream_ptr(il)),
               "nnd_gev_squeeze_soft_argmin")
    ref_cost = oi.squeeze_cost(geo.reshape(-1, D), (B, G, H, W1, D), weight, bias)
    scale = np.abs(ref_cost).max()
    np.testing.assert_allclose(cost.cpu().numpy(), ref_cost, rtol=VOLUME_RTOL, atol=VOLUME_RTOL * scale)
    ref = oi.squeeze_soft_argmin(geo.reshape(-1, D), (B, G, H, W1, D), weight, bias)
    np.testing.assert_allclose(disp.cpu().numpy(), ref, rtol=0, atol=SOFTARGMIN_ATOL * max(1.0, D / 160))


def test_squeeze_soft_argmin_reference_layout_fallback():
    """W2 % 8 != 0 keeps the reference row layout: cuDNN Conv3d + the soft-argmin kernel, same result."""
    import nndepth_b200 as nb
    rng = np.random.default_rng(6)
    B, G, H, W1, D = 1, 8, 3, 8, 20
    geo = rng.standard_normal((B, G, H, W1, D), dtype=np.float32)
    weight = (rng.standard_normal((1, G, 3, 3, 3)) * 0.3).astype(np.float32)
    pyr = oi.volume_pyramids(geo, geo.transpose(0, 1, 4, 2, 3), 2)[1]
    cv = nb.GeometryAwareCostVolume.from_pyramids(pyr[:2], pyr[:2], B, H, 2, 4, G)
    assert not cv._interleaved
    sq = torch.nn.Conv3d(G, 1, 3, 1, 1).cuda()
    sq.weight.data.copy_(dev(weight))
    with torch.backends.cudnn.flags(allow_tf32=False):
        disp = cv.init_disparity(sq)
    ref = oi.squeeze_soft_argmin(pyr[0], (B, G, H, W1, D), weight, sq.bias.detach().cpu().numpy())
    np.testing.assert_allclose(disp.cpu().numpy(), ref, rtol=0, atol=SOFTARGMIN_ATOL)


def test_squeeze_rejects_other_convolutions():
    import nndepth_b200 as nb
    geo = np.zeros((1, 8, 2, 8, 16), dtype=np.float32)
    pyr = oi.volume_pyramids(geo, geo.transpose(0, 1, 4, 2, 3), 1)[1]
    cv = nb.GeometryAwareCostVolume.from_pyramids(pyr[:1], pyr[:1], 1, 2, 1, 4, 8)
    with pytest.raises(RuntimeError):
        cv.init_disparity(torch.nn.Conv3d(8, 2, 3, 1, 1).cuda())
    with pytest.raises(RuntimeError):
        cv.init_disparity(torch.nn.Conv3d(8, 1, 3, 1, 0).cuda())

"""Pin oracle/igev.py to the reference outputs in tests/golden/igev.npz."""
import numpy as np

from oracle import igev as oi

REGIMES = ["int", "sub", "oob"]


def test_groupwise_volume_uses_first_g_squared_channels(golden):
    g = golden("igev")
    G = int(g["num_groups"])
    np.testing.assert_array_equal(g["feat_volume"], g["feat_volume_first64_only"])
    vol = oi.groupwise_volume(g["fmap1"], g["fmap2"], G)
    ref = g["feat_volume"]
    np.testing.assert_allclose(vol, ref, rtol=1e-5, atol=1e-5 * np.abs(ref).max())
    f1 = g["fmap1"].copy()
    f1[:, G * G:] = 7.0
    np.testing.assert_array_equal(oi.groupwise_volume(f1, g["fmap2"], G), vol)


def test_pyramids_bit_exact(golden):
    g = golden("igev")
    fp, gp = oi.volume_pyramids(g["feat_volume"], g["geo_volume"], 4)
    for lvl in range(5):
        np.testing.assert_array_equal(fp[lvl], g[f"feat_pyr{lvl}"])
        np.testing.assert_array_equal(gp[lvl], g[f"geo_pyr{lvl}"])


def test_dual_lookup_bit_exact(golden):
    g = golden("igev")
    G = int(g["num_groups"])
    fp = [g[f"feat_pyr{lvl}"] for lvl in range(5)]
    gp = [g[f"geo_pyr{lvl}"] for lvl in range(5)]
    for regime in REGIMES:
        out = oi.gev_lookup(fp, gp, g[f"coords_{regime}"], 4, 4, G)
        assert out.shape == g[f"out_{regime}"].shape == (2, 4 * 2 * G * 9, 3, 24)
        np.testing.assert_array_equal(out, g[f"out_{regime}"])


def test_soft_argmin(golden):
    g = golden("igev")
    z = g["sa_logits"]
    p = oi.softmax_disparity(z)
    np.testing.assert_allclose(p, g["sa_softmax"], rtol=2e-6, atol=1e-9)
    np.testing.assert_allclose(oi.regress_disparity(g["sa_softmax"], 24), g["sa_disp"], rtol=1e-6, atol=1e-6)
    fused = oi.soft_argmin(z)
    assert fused.shape == g["sa_disp"].shape
    np.testing.assert_allclose(fused, g["sa_disp"], rtol=1e-5, atol=1e-5)
    assert fused[0, 0, 0, 1] == -5.0                     # one-hot row
    np.testing.assert_allclose(fused[0, 0, 0, 0], -11.5, rtol=1e-6)   # uniform row: mean of 0..23


def test_squeeze_soft_argmin(golden):
    """cv_squeezer (Conv3d) + softmax + regress_disparity against the reference model's own op chain."""
    g = golden("igev_squeeze")
    cost = oi.squeeze_cost(g["geo_pyr0"], g["shape"], g["weight"], g["bias"])
    assert cost.shape == g["cost"].shape
    np.testing.assert_allclose(cost, g["cost"], rtol=1e-5, atol=1e-5 * np.abs(g["cost"]).max())
    disp = oi.squeeze_soft_argmin(g["geo_pyr0"], g["shape"], g["weight"], g["bias"])
    np.testing.assert_allclose(disp, g["disp"], rtol=1e-4, atol=1e-4)
    # the padding is zero padding: a constant volume gives smaller sums on the faces than inside
    B, G, H, W1, W2 = (int(v) for v in g["shape"])
    ones = np.ones((B * G * H * W1, W2), dtype=np.float32)
    c1 = oi.squeeze_cost(ones, g["shape"], np.ones((1, G, 3, 3, 3), np.float32))
    assert c1[0, 1, 1, 1] == 27 * G and c1[0, 0, 0, 0] == 8 * G and c1[0, 0, 1, 1] == 18 * G

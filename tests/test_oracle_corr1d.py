"""Pin the numpy oracle (oracle/corr1d.py) to outputs of the unmodified reference (tests/golden)."""
import numpy as np
import pytest

from oracle import corr1d as oc

CASES = ["corr1d_small", "corr1d_odd", "corr1d_kitti_row", "corr1d_r3l3"]
REGIMES = ["int", "sub", "oob"]


def test_linear_sampler_known_answers(golden):
    g = golden("sampler_kats")
    np.testing.assert_array_equal(oc.linear_sampler(g["kat_row"], g["kat_x"]), g["kat_out"])
    np.testing.assert_array_equal(g["kat_out"][0], np.float32([0, 0, 0, 0.25, 8.75, 9, 9, 9]))
    np.testing.assert_array_equal(oc.avg_pool_pairs(g["pool_in"]), g["pool_out"])
    np.testing.assert_array_equal(g["pool_out"][0], np.float32([0.5, 2.5, 4.5]))


@pytest.mark.parametrize("w2", [240, 160, 156, 120, 80, 78, 60, 40, 39, 30, 20, 19, 9, 5, 2])
def test_integer_round_trip_table(golden, w2):
    """fact 6: x/(w2-1)*(w2-1) does not round-trip integers; floor/ceil must match ATen bit for bit."""
    g = golden("sampler_kats")
    x = np.arange(w2, dtype=np.float32)[None]
    _, i0, i1 = oc.sampler_indices(x, w2)
    np.testing.assert_array_equal(i0[0], g[f"aten_rt_floor_{w2}"])
    np.testing.assert_array_equal(i1[0], g[f"aten_rt_ceil_{w2}"])
    ramp = (np.arange(w2, dtype=np.float32) * np.float32(1.5) - np.float32(3))[None]
    np.testing.assert_array_equal(oc.linear_sampler(ramp, x)[0], g[f"rt_out_{w2}"])


def test_round_trip_is_not_identity(golden):
    g = golden("sampler_kats")
    moved = (g["aten_rt_floor_160"] != np.arange(160)).sum() + (g["aten_rt_ceil_160"] != np.arange(160)).sum()
    assert moved == 27


@pytest.mark.parametrize("case", CASES)
def test_volume_and_pyramid(golden, case):
    g = golden(case)
    L = int(g["num_levels"])
    vol = oc.all_pairs_correlation(g["fmap1"], g["fmap2"])
    ref0 = g["pyr0"]
    # fp32 tolerance of BASELINE.json: 1e-5 relative (to the volume's scale: values cross zero)
    scale = np.abs(ref0).max()
    np.testing.assert_allclose(vol.reshape(ref0.shape), ref0, rtol=1e-5, atol=1e-5 * scale)
    # pooling is exact arithmetic given the same level-0 input: bit-exact
    pyr = oc.build_pyramid(ref0, L)
    assert len(pyr) == L + 1
    for lvl in range(L + 1):
        np.testing.assert_array_equal(pyr[lvl], g[f"pyr{lvl}"])


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("regime", REGIMES)
def test_lookup_bit_exact_on_reference_pyramid(golden, case, regime):
    g = golden(case)
    L, r = int(g["num_levels"]), int(g["radius"])
    pyr = [g[f"pyr{lvl}"] for lvl in range(L + 1)]
    out = oc.lookup(pyr, g[f"coords_{regime}"], L, r)
    assert out.dtype == np.float32 and out.shape == g[f"out_{regime}"].shape
    np.testing.assert_array_equal(out, g[f"out_{regime}"])
    widths = [p.shape[1] for p in pyr]
    for lvl, (i0, i1) in enumerate(oc.lookup_indices(widths, g[f"coords_{regime}"], L, r)):
        np.testing.assert_array_equal(i0, g[f"aten_i0_{regime}_{lvl}"])
        np.testing.assert_array_equal(i1, g[f"aten_i1_{regime}_{lvl}"])


@pytest.mark.parametrize("case", CASES)
def test_block_end_to_end(golden, case):
    g = golden(case)
    L, r = int(g["num_levels"]), int(g["radius"])
    blk = oc.CorrBlock1D(g["fmap1"], g["fmap2"], L, r)
    for regime in REGIMES:
        out = blk(g[f"coords_{regime}"])
        ref = g[f"out_{regime}"]
        np.testing.assert_allclose(out, ref, rtol=1e-5, atol=1e-5 * np.abs(ref).max())


def test_group_block(golden):
    g = golden("group_corr1d")
    G = int(g["num_groups"])
    vol = oc.group_all_pairs_correlation(g["fmap1"], g["fmap2"], G)
    ref0 = g["pyr0"]
    np.testing.assert_allclose(vol.reshape(ref0.shape), ref0, rtol=1e-5, atol=1e-5 * np.abs(ref0).max())
    pyr = [g[f"pyr{lvl}"] for lvl in range(5)]
    for regime in REGIMES:
        out = oc.group_lookup(pyr, g[f"coords_{regime}"], 4, 4, G)
        np.testing.assert_array_equal(out, g[f"out_{regime}"])


def test_level_width_one_is_rejected():
    with pytest.raises(ValueError):
        oc.sampler_indices(np.zeros((1, 1), np.float32), 1)

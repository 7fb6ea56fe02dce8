"""Pin oracle/agcl.py to the reference outputs in tests/golden/agcl.npz."""
import numpy as np
import pytest

from oracle import agcl as oa

# mean over C/4 channels is a tree/vector reduction inside ATen: order differs -> fp32 tolerance
TOL = dict(rtol=1e-5, atol=2e-6)


def test_warp_matches_reference(golden):
    g = golden("agcl")
    N, C, H, W = g["fmap2"].shape
    coords = (oa.coords_grid(N, H, W) + g["flow"]).transpose(0, 2, 3, 1)
    np.testing.assert_array_equal(oa.bilinear_sampler(g["fmap2"], coords), g["warped_right"])
    assert not g["warped_right"][0, :, 0, 0].any()       # flow (-30, 4): all corners outside


@pytest.mark.parametrize("small", [False, True])
def test_iter_mode(golden, small):
    g = golden("agcl")
    out = oa.corr_iter(g["fmap1"], g["fmap2"], g["flow"], small)
    ref = g["iter_3x3" if small else "iter_1x9"]
    assert out.shape == ref.shape == (2, 36, 6, 10)
    np.testing.assert_allclose(out, ref, **TOL)


@pytest.mark.parametrize("small", [False, True])
def test_offset_mode(golden, small):
    g = golden("agcl")
    out = oa.corr_att_offset(g["fmap1"], g["fmap2"], g["flow"], g["extra_offset"], small)
    ref = g["offset_3x3" if small else "offset_1x9"]
    np.testing.assert_allclose(out, ref, **TOL)


def test_offset_mode_with_attention_hook(golden):
    g = golden("agcl")

    def att(left, right):
        return left * np.float32(0.5) + right[:, ::-1] * np.float32(0.25), right - left * np.float32(0.125)

    out = oa.AGCL(g["fmap1"], g["fmap2"], att=att)(g["flow"], g["extra_offset"], False, False)
    np.testing.assert_allclose(out, g["offset_att_1x9"], **TOL)

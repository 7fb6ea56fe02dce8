"""GPU parity: CREStereo AGCL (offset mode with deformable sampling, iter mode) vs reference goldens/oracle."""
import numpy as np
import pytest
import torch

from oracle import agcl as oa

pytestmark = pytest.mark.gpu

# mean over C/4 channels: summation order differs from ATen's -> fp32 tolerance (BASELINE: 1e-5 relative)
TOL = dict(rtol=1e-5, atol=2e-6)


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("small", [False, True])
def test_iter_mode_golden(golden, small):
    import nndepth_b200 as nb
    g = golden("agcl")
    agcl = nb.AGCL(dev(g["fmap1"]), dev(g["fmap2"]))
    out = agcl(dev(g["flow"]), None, small_patch=small, iter_mode=True).cpu().numpy()
    ref = g["iter_3x3" if small else "iter_1x9"]
    assert out.shape == ref.shape == (2, 36, 6, 10)
    np.testing.assert_allclose(out, ref, **TOL)


@pytest.mark.parametrize("small", [False, True])
def test_offset_mode_golden(golden, small):
    import nndepth_b200 as nb
    g = golden("agcl")
    agcl = nb.AGCL(dev(g["fmap1"]), dev(g["fmap2"]))
    out = agcl(dev(g["flow"]), dev(g["extra_offset"]), small_patch=small, iter_mode=False).cpu().numpy()
    np.testing.assert_allclose(out, g["offset_3x3" if small else "offset_1x9"], **TOL)


def test_offset_mode_with_attention_hook(golden):
    import nndepth_b200 as nb
    g = golden("agcl")
    calls = []

    def att(left, right):
        calls.append(1)
        return left * 0.5 + right.flip(1) * 0.25, right - left * 0.125

    agcl = nb.AGCL(dev(g["fmap1"]), dev(g["fmap2"]), att=att)
    flow, offs = dev(g["flow"]), dev(g["extra_offset"])
    out = agcl(flow, offs, small_patch=False, iter_mode=False).cpu().numpy()
    np.testing.assert_allclose(out, g["offset_att_1x9"], **TOL)
    agcl(flow, offs, small_patch=True, iter_mode=False)
    assert len(calls) == 1      # attention output is a pure function of the maps: cached


@pytest.mark.parametrize("shape", [(2, 64, 22, 40), (1, 32, 45, 80), (1, 8, 7, 61), (1, 4, 2, 2),
                                   (1, 16, 4, 6), (1, 48, 5, 9), (1, 512, 3, 4), (1, 528, 2, 3), (2, 256, 5, 7), (1, 128, 3, 11)])
@pytest.mark.parametrize("small", [False, True])
def test_against_oracle(shape, small):
    """Scales of BASELINE config 3 (1/32 and 1/16 of 720x1280, reduced channels) + ragged shapes.

    C % 16 == 0 and C <= 512 take the channels-last kernels (1, 2, 3 and 4 chunks per lane; idle sub-lanes
    at C = 16 / 48; C = 128 and 256 the four-pixels-per-warp variant), the others (C = 8, 4, 528) the generic
    NCHW kernels."""
    import nndepth_b200 as nb
    rng = np.random.default_rng(7)
    N, C, H, W = shape
    f1 = rng.standard_normal(shape, dtype=np.float32)
    f2 = rng.standard_normal(shape, dtype=np.float32)
    flow = (rng.standard_normal((N, 2, H, W)) * 3).astype(np.float32)
    flow[0, :, 0, 0] = (-1000.0, 3.0)
    flow[0, :, -1, -1] = (0.0, 0.0)
    offs = rng.uniform(-1, 1, size=(N, 18, H, W)).astype(np.float32)
    agcl = nb.AGCL(dev(f1), dev(f2))
    got_it = agcl(dev(flow), None, small_patch=small, iter_mode=True).cpu().numpy()
    got_of = agcl(dev(flow), dev(offs), small_patch=small, iter_mode=False).cpu().numpy()
    np.testing.assert_allclose(got_it, oa.corr_iter(f1, f2, flow, small), **TOL)
    np.testing.assert_allclose(got_of, oa.corr_att_offset(f1, f2, flow, offs, small), **TOL)


def test_config3_full_size_properties():
    """BASELINE config 3 finest scale (N4, 256 ch, 90x160): properties that need no CPU oracle."""
    import nndepth_b200 as nb
    torch.manual_seed(3)
    N, C, H, W = 4, 256, 90, 160
    f1 = torch.randn(N, C, H, W, device="cuda")
    f2 = torch.randn(N, C, H, W, device="cuda")
    agcl = nb.AGCL(f1, f2)
    zero = torch.zeros(N, 2, H, W, device="cuda")
    # zero flow + zero offsets: both modes sample the right map at integer positions.  Offset mode
    # zero-pads, iter mode replicate-pads, so they agree away from the border ...
    it = agcl(zero, None, small_patch=False, iter_mode=True)
    of = agcl(zero, torch.zeros(N, 18, H, W, device="cuda"), small_patch=False, iter_mode=False)
    assert torch.allclose(it[..., 4:-4], of[..., 4:-4], rtol=1e-4, atol=1e-5)
    # ... and the centre tap (k = 4) is the plain group-wise mean of f1 * f2
    ref = (f1 * f2).reshape(N, 4, C // 4, H, W).double().mean(2).float()
    assert torch.allclose(it[:, 4::9], ref, rtol=1e-4, atol=1e-5)
    # linearity in the left map
    g1 = torch.randn_like(f1)
    flow = torch.randn(N, 2, H, W, device="cuda") * 3
    a = nb.AGCL(f1, f2)(flow, None, True, True)
    b = nb.AGCL(g1, f2)(flow, None, True, True)
    c = nb.AGCL(2 * f1 + g1, f2)(flow, None, True, True)
    assert torch.allclose(c, 2 * a + b, rtol=1e-4, atol=1e-5)
    # a flow pointing far outside the image gives exactly zero correlation in both modes
    far = torch.full((N, 2, H, W), 1e4, device="cuda")
    assert not agcl(far, None, False, True).any()
    assert not agcl(far, torch.zeros(N, 18, H, W, device="cuda"), True, False).any()


def test_errors():
    import nndepth_b200 as nb
    f = torch.randn(1, 30, 4, 4, device="cuda")
    with pytest.raises(nb.NNDepthError, match="divisible"):
        nb.AGCL(f, f)(torch.zeros(1, 2, 4, 4, device="cuda"), None, False, True)
    f = torch.randn(1, 8, 4, 4, device="cuda")
    with pytest.raises(RuntimeError, match="extra_offset"):
        nb.AGCL(f, f)(torch.zeros(1, 2, 4, 4, device="cuda"), torch.zeros(1, 9, 4, 4, device="cuda"))
    with pytest.raises(RuntimeError, match="flow"):
        nb.AGCL(f, f)(torch.zeros(1, 1, 4, 4, device="cuda"), None, False, True)

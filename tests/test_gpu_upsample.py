"""GPU parity of the fused convex upsampling vs the reference's torch chain (raft_stereo/model.py:93-105)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def reference_chain(flow, mask, rate):
    """Verbatim op chain of the reference (softmax / unfold / mul / sum / permute / reshape), on the device."""
    import torch.nn.functional as F
    N, _, H, W = flow.shape
    mask = mask.view(N, 1, 9, rate, rate, H, W)
    mask = torch.softmax(mask, dim=2)
    up = F.unfold(rate * flow, (3, 3), padding=1).view(N, 1, 9, 1, 1, H, W)
    up = torch.sum(mask * up, dim=2).permute(0, 1, 4, 2, 5, 3)
    return up.reshape(N, 1, rate * H, rate * W)


@pytest.mark.parametrize("shape,rate", [((2, 6, 10), 8), ((1, 48, 156), 8), ((3, 5, 33), 4), ((1, 1, 1), 8), ((2, 7, 9), 2)])
def test_matches_reference_chain(shape, rate):
    import nndepth_b200 as nb
    N, H, W = shape
    torch.manual_seed(N * H + W)
    flow = torch.randn(N, 1, H, W, device="cuda") * 20
    mask = torch.randn(N, 9 * rate * rate, H, W, device="cuda") * 3
    got = nb.convex_upsample(flow, mask, rate)
    ref = reference_chain(flow.double(), mask.double(), rate).float()
    assert got.shape == ref.shape == (N, 1, rate * H, rate * W)
    assert (got - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item())
    if rate == 8:   # channels-last mask: same numbers, consumed without a layout copy
        cl = mask.contiguous(memory_format=torch.channels_last)
        got_cl = nb.convex_upsample(flow, cl, rate)
        assert (got_cl - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item())
    # a per-channel bias folded into the kernel == adding it to the mask first
    bias = torch.randn(9 * rate * rate, device="cuda")
    with_bias = nb.convex_upsample(flow, mask, rate, mask_bias=bias)
    ref_b = reference_chain(flow.double(), (mask + bias.view(1, -1, 1, 1)).double(), rate).float()
    assert (with_bias - ref_b).abs().max().item() <= 1e-5 * max(1.0, ref_b.abs().max().item())
    if rate == 8:
        with_bias_cl = nb.convex_upsample(flow, mask.contiguous(memory_format=torch.channels_last), rate, mask_bias=bias)
        assert (with_bias_cl - ref_b).abs().max().item() <= 1e-5 * max(1.0, ref_b.abs().max().item())
    # the folded 0.25 scale is exact (power of two): identical bits to scaling the mask first
    assert torch.equal(nb.convex_upsample(flow, mask, rate, mask_scale=0.25), nb.convex_upsample(flow, 0.25 * mask, rate))


@pytest.mark.parametrize("shape", [(2, 6, 10), (1, 48, 156), (1, 1, 1)])
def test_fp16_channels_last_mask(shape):
    """fp16 logits (mask head run as fp16 convolutions): the kernel converts exactly, so the result equals the
    fp32 path on the same (fp16-representable) logits."""
    import nndepth_b200 as nb
    N, H, W = shape
    torch.manual_seed(3)
    flow = torch.randn(N, 1, H, W, device="cuda") * 20
    mask16 = (torch.randn(N, 576, H, W, device="cuda") * 3).half().contiguous(memory_format=torch.channels_last)
    bias = torch.randn(576, device="cuda")
    got = nb.convex_upsample(flow, mask16, 8, mask_scale=0.25, mask_bias=bias)
    same = nb.convex_upsample(flow, mask16.float(), 8, mask_scale=0.25, mask_bias=bias)
    assert torch.equal(got, same)
    ref = reference_chain(flow.double(), 0.25 * (mask16.double() + bias.double().view(1, -1, 1, 1)), 8).float()
    assert (got - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item())


def test_convexity_and_constant_field():
    """A convex combination of a constant field is that constant (times rate), whatever the mask."""
    import nndepth_b200 as nb
    flow = torch.full((1, 1, 9, 11), -3.5, device="cuda")
    mask = torch.randn(1, 576, 9, 11, device="cuda") * 5
    out = nb.convex_upsample(flow, mask, 8)
    inner = out[..., 8:-8, 8:-8]                      # away from the zero-padded border
    assert (inner + 28.0).abs().max().item() <= 1e-5 * 28
    assert out.min().item() >= -28.0 - 1e-4 and out.max().item() <= 1e-6     # border mixes in zeros, never overshoots


def test_errors():
    import nndepth_b200 as nb
    flow = torch.zeros(1, 1, 4, 4, device="cuda")
    with pytest.raises(RuntimeError, match="mask"):
        nb.convex_upsample(flow, torch.zeros(1, 9 * 64, 4, 5, device="cuda"), 8)
    with pytest.raises(nb.NNDepthError, match="rate"):
        nb.convex_upsample(flow, torch.zeros(1, 81, 4, 4, device="cuda"), 3)
    with pytest.raises(RuntimeError, match="CUDA"):
        nb.convex_upsample(flow.cpu(), torch.zeros(1, 576, 4, 4), 8)

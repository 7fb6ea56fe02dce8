"""GPU parity: the single-flow-channel convolutions of the update block against PyTorch fp32 convolutions
(reference blocks/update_block.py:23,36,53,60 are plain nn.Conv2d; tolerance = fp32 summation-order noise)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,cout", [((2, 1, 9, 21), 128), ((1, 1, 48, 156), 128), ((1, 1, 3, 5), 64), ((3, 1, 7, 8), 32)])
def test_flow_conv7x7_relu(shape, cout):
    from nndepth_b200.raft_stereo import flow_conv7x7_relu
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(1, cout, 7, padding=3).cuda()
    flow = torch.randn(*shape, device="cuda") * 5
    with torch.no_grad(), torch.backends.cudnn.flags(allow_tf32=False):
        ref = F.relu(conv(flow))
        got = flow_conv7x7_relu(conv, flow)
    assert got.shape == ref.shape and got.permute(0, 2, 3, 1).is_contiguous()
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-5)
    assert (got == 0).any() and (got > 0).any()


def test_flow_conv7x7_fp16_output_and_cat():
    from nndepth_b200.raft_stereo import flow_conv7x7_relu, nhwc_cat_f16
    torch.manual_seed(4)
    conv = torch.nn.Conv2d(1, 128, 7, padding=3).cuda()
    flow = torch.randn(2, 1, 9, 21, device="cuda") * 5
    with torch.no_grad():
        f32 = flow_conv7x7_relu(conv, flow)
        f16 = flow_conv7x7_relu(conv, flow, half=True)
        assert f16.dtype == torch.float16 and torch.equal(f16, f32.half())
        a = torch.randn(2, 192, 9, 21, device="cuda").contiguous(memory_format=torch.channels_last)
        b = torch.randn(2, 64, 9, 21, device="cuda").half().contiguous(memory_format=torch.channels_last)
        cat = nhwc_cat_f16(a, b)
        assert cat.dtype == torch.float16 and cat.permute(0, 2, 3, 1).is_contiguous()
        assert torch.equal(cat, torch.cat([a.half(), b], 1))
        assert torch.equal(nhwc_cat_f16(b, a), torch.cat([b, a.half()], 1))


def test_flow_conv7x7_falls_back_for_two_flow_channels():
    from nndepth_b200.raft_stereo import flow_conv7x7_relu
    conv = torch.nn.Conv2d(2, 128, 7, padding=3).cuda()
    flow = torch.randn(1, 2, 6, 9, device="cuda")
    with torch.no_grad(), torch.backends.cudnn.flags(allow_tf32=False):
        torch.testing.assert_close(flow_conv7x7_relu(conv, flow), F.relu(conv(flow)), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("shape", [(2, 128, 9, 21), (1, 128, 48, 156), (1, 256, 9, 13), (1, 512, 5, 7), (1, 256, 1, 1)])
@pytest.mark.parametrize("bias", [True, False])
def test_flow_head_tail(shape, bias):
    from nndepth_b200.raft_stereo import flow_head_tail
    torch.manual_seed(1)
    N, C, H, W = shape
    conv = torch.nn.Conv2d(C, 1, 3, padding=1, bias=bias).cuda().to(memory_format=torch.channels_last)
    x = torch.randn(N, C, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
    coords = torch.randn(N, 1, H, W, device="cuda") * 30
    org = torch.arange(W, device="cuda").float().view(1, 1, 1, W).repeat(N, 1, H, 1)
    with torch.no_grad(), torch.backends.cudnn.flags(allow_tf32=False):
        ref = conv(x)
        delta = flow_head_tail(conv, x)
        new_coords, new_flow = flow_head_tail(conv, x, coords, org)
    scale = ref.abs().max().item()
    torch.testing.assert_close(delta, ref, rtol=1e-5, atol=1e-5 * scale)
    # the fused update is the loop's own two fp32 operations on the same delta: bit-exact
    assert torch.equal(new_coords, coords + delta)
    assert torch.equal(new_flow, (coords + delta) - org)


@pytest.mark.parametrize("C", [128, 256])
def test_flow_head_tail_fp16_input(C):
    """fp16 channels-last input (the flow head's first convolution run in fp16): exact conversion, fp32 arithmetic."""
    from nndepth_b200.raft_stereo import flow_head_tail
    torch.manual_seed(2)
    conv = torch.nn.Conv2d(C, 1, 3, padding=1).cuda()
    x16 = torch.randn(2, C, 7, 19, device="cuda").half().contiguous(memory_format=torch.channels_last)
    with torch.no_grad(), torch.backends.cudnn.flags(allow_tf32=False):
        ref = conv(x16.float())
        delta = flow_head_tail(conv, x16)
        same = flow_head_tail(conv, x16.float())
    assert torch.equal(delta, same)
    torch.testing.assert_close(delta, ref, rtol=1e-5, atol=1e-5 * ref.abs().max().item())


def test_flow_head_tail_declines_other_shapes():
    from nndepth_b200.raft_stereo import flow_head_tail
    x = torch.randn(1, 128, 4, 4, device="cuda").contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        x64 = torch.randn(1, 64, 4, 4, device="cuda").contiguous(memory_format=torch.channels_last)
        assert flow_head_tail(torch.nn.Conv2d(64, 1, 3, padding=1).cuda(), x64) is None        # C = 64
        x256 = torch.randn(1, 256, 4, 4, device="cuda")                                       # NCHW, not channels-last
        assert flow_head_tail(torch.nn.Conv2d(256, 1, 3, padding=1).cuda(), x256) is None
        assert flow_head_tail(torch.nn.Conv2d(256, 2, 3, padding=1).cuda(),
                              x256.contiguous(memory_format=torch.channels_last)) is None

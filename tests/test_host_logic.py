"""CPU tests of the host-side helpers around the kernels (no GPU needed): TF32 rounding, BatchNorm folding, the
weight split of the ConvGRU, the model shell's state-dict compatibility with the reference layout."""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from nndepth_b200 import raft_stereo as rs


def test_rn_tf32_is_round_to_nearest_on_10_mantissa_bits():
    x = torch.tensor([1.0, 1.0 + 2 ** -11, 1.0 + 2 ** -11 + 2 ** -20, 1.0 + 2 ** -10, -3.1415927, 1e-30, 65504.0, 0.0])
    y = rs.rn_tf32(x)
    bits = y.view(torch.int32)
    assert torch.all((bits & 0x1FFF) == 0)                       # low 13 mantissa bits cleared
    assert torch.all((y - x).abs() <= x.abs() * 2.0 ** -11)       # half an ulp of a 10-bit mantissa
    assert y[0] == 1.0 and y[2] == 1.0 + 2 ** -10 and y[3] == 1.0 + 2 ** -10
    # ties round away from zero (cvt.rna), like the device kernels
    assert y[1] == 1.0 + 2 ** -10


def test_fold_bn_matches_conv_followed_by_eval_batchnorm():
    torch.manual_seed(0)
    conv = nn.Conv2d(5, 7, 3, padding=1)
    bn = nn.BatchNorm2d(7).eval()
    with torch.no_grad():
        bn.running_mean.uniform_(-1, 1)
        bn.running_var.uniform_(0.5, 2)
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-1, 1)
    x = torch.randn(2, 5, 6, 8)
    w, b = rs.fold_bn(conv, bn)
    with torch.no_grad():
        ref = bn(conv(x))
        got = F.conv2d(x, w, b, padding=1)
    assert torch.allclose(got, ref, rtol=1e-5, atol=1e-5)
    w32, _ = rs.fold_bn(conv, bn, tf32=True)
    assert torch.equal(w32, rs.rn_tf32(w))


def test_gru_weight_split_reconstructs_the_weights_exactly():
    torch.manual_seed(1)
    gru = rs.SepConvGRU(hidden_dim=8, input_dim=12)
    (wzr, bzr), (wq, bq), pad = gru._split_weights("1")
    cin = 8 + 12
    assert wzr.shape == (16, 2 * cin, 1, 5) and wq.shape == (8, 2 * cin, 1, 5) and pad == (0, 2)
    full = torch.cat([gru.convz1.weight, gru.convr1.weight], 0)
    assert torch.equal(wzr[:, :cin] + wzr[:, cin:], full)            # w_hi + w_lo == w, exactly
    assert torch.equal(wzr[:, :cin], rs.rn_tf32(full.detach()))
    assert torch.equal(bzr, torch.cat([gru.convz1.bias, gru.convr1.bias], 0))
    # the torch-op form of the weight-split recurrence equals the plain recurrence up to TF32 rounding of the activations
    h, x = torch.tanh(torch.randn(1, 8, 4, 6)), torch.randn(1, 12, 4, 6)
    with torch.no_grad():
        ref = gru._half_step(h, x, "1")
        got = gru._half_step_wsplit(h, x, "1")
    assert (got - ref).abs().max().item() < 5e-3


def test_model_shell_parameter_names_follow_the_reference_layout():
    torch.manual_seed(0)
    model = rs.BaseRAFTStereo(iters=2)
    names = set(model.state_dict().keys())
    for expected in ("fnet.conv1.weight", "fnet.layer1.0.conv1.weight", "fnet.layer3.1.downsample.0.weight", "fnet.conv2.bias",
                     "cnet_proj.0.weight", "update_block.encoder.convc1.weight", "update_block.gru.convz1.weight",
                     "update_block.gru.convq2.bias", "update_block.flow_head.conv2.weight", "update_block.mask.2.weight"):
        assert expected in names, expected
    assert model.update_block.encoder.convc1.weight.shape == (256, 36, 1, 1)
    assert model.update_block.mask[2].weight.shape == (576, 256, 1, 1)
    # CPU forward of the shell with the oracle correlation block (the path bench.py --impl reference times)
    from oracle import torch_port
    model.corr_fn = torch_port.CorrBlock1D
    model.eval()
    with torch.no_grad():
        outs = model(torch.rand(1, 3, 64, 160) * 2 - 1, torch.rand(1, 3, 64, 160) * 2 - 1)
    assert len(outs) == 2 and outs[-1]["up_disp"].shape == (1, 1, 64, 160) and torch.isfinite(outs[-1]["up_disp"]).all()
    assert np.isfinite(outs[0]["up_disp"].numpy()).all()


def test_gru_split_weights_reconstruct_fp32_weights():
    """[w_hi; w_lo] of the weight-split ConvGRU: hi + lo reproduces the fp32 weight to 2^-22 (TF32 split) and to about
    2^-17 relative (fp16 split, where w_lo runs into fp16's subnormal spacing) -- versus 2^-11 for a plain rounding."""
    import torch
    from nndepth_b200.raft_stereo import SepConvGRU
    torch.manual_seed(0)
    gru = SepConvGRU(hidden_dim=16, input_dim=24).eval()
    for tag in "12":
        wz, wr, wq = (getattr(gru, f"conv{g}{tag}").weight.detach() for g in "zrq")
        for half, tol in ((False, 2.0 ** -21), (True, 2.0 ** -15)):
            (wzr, bzr), (wq2, bq), pad = gru._split_weights(tag, half=half)
            cin = wz.shape[1]
            assert wzr.shape[1] == 2 * cin and wq2.shape[1] == 2 * cin and pad == getattr(gru, f"convz{tag}").padding
            assert wzr.dtype == (torch.float16 if half else torch.float32)
            rec_zr = wzr[:, :cin].float() + wzr[:, cin:].float()
            rec_q = wq2[:, :cin].float() + wq2[:, cin:].float()
            ref_zr = torch.cat([wz, wr], 0)
            scale = ref_zr.abs().max().item()
            assert (rec_zr - ref_zr).abs().max().item() <= tol * scale
            assert (rec_q - wq).abs().max().item() <= tol * scale
            # the plain rounding alone is 2^-12 .. 2^-11 off: the low part matters
            assert (wzr[:, :cin].float() - ref_zr).abs().max().item() > 2.0 ** -14 * scale
            assert bzr.dtype == torch.float32 and bq.dtype == torch.float32


def test_fp16_rounding_agrees_with_rn_tf32():
    import torch
    from nndepth_b200.raft_stereo import rn_tf32
    x = torch.tensor([1.0, 1.0 + 2.0 ** -11, 1.0 + 2.0 ** -10, 1.0 + 3 * 2.0 ** -12, -3.1415927, 1e-30, 65504.0])
    y = rn_tf32(x)
    assert torch.equal(y.view(torch.int32) & 0x1FFF, torch.zeros_like(y, dtype=torch.int32))     # low 13 bits cleared
    assert (y - x).abs().max().item() <= (x.abs() * 2.0 ** -11).max().item()
    assert y[0] == 1.0 and y[2] == 1.0 + 2.0 ** -10
    # fp16 carries the same mantissa: inside fp16's normal range the two roundings agree (ties aside)
    z = torch.randn(4096) * 3
    same = (rn_tf32(z) == z.half().float()).float().mean().item()
    assert same > 0.999


def test_weight_caches_follow_in_place_updates():
    """Split / folded weight caches are keyed on the parameters' versions: loading a checkpoint after a first forward
    must not leave stale tensors behind."""
    import torch
    from nndepth_b200.raft_stereo import ResidualBlock, SepConvGRU
    torch.manual_seed(2)
    gru = SepConvGRU(hidden_dim=8, input_dim=8).eval()
    (w_a, _), _, _ = gru._split_weights("1")
    assert gru._split_weights("1")[0][0] is w_a                      # cached while nothing changes
    with torch.no_grad():
        gru.convz1.weight.mul_(2.0)
    (w_b, _), _, _ = gru._split_weights("1")
    cin = gru.convz1.weight.shape[1]
    assert torch.allclose(w_b[:8, :cin] + w_b[:8, cin:], gru.convz1.weight.detach(), atol=1e-7)
    assert not torch.equal(w_a, w_b)
    block = ResidualBlock(4, 4, norm_fn="batch").eval()
    f_a = block._folded()[0][0]
    with torch.no_grad():
        block.norm1.running_var.fill_(4.0)
    f_b = block._folded()[0][0]
    assert not torch.equal(f_a, f_b)


def test_fused_gate_weights_follow_parameter_updates_and_keep_autograd():
    """the concatenated z|r gate weights are keyed on the parameters' versions (no stale snapshot after
    load_state_dict / in-place updates) and are bypassed under grad mode (convz / convr keep their gradients)."""
    import torch
    from nndepth_b200.raft_stereo import SepConvGRU
    torch.manual_seed(0)
    gru = SepConvGRU(hidden_dim=8, input_dim=12).eval()
    plain = SepConvGRU(hidden_dim=8, input_dim=12).eval()
    h, x = torch.tanh(torch.randn(1, 8, 5, 7)), torch.randn(1, 12, 5, 7)
    gru.fuse_gates()
    with torch.no_grad():
        gru(h, x)                                           # populates the fused cache with the OLD weights
        gru.load_state_dict(plain.state_dict())             # in-place update
        torch.testing.assert_close(gru(h, x), plain(h, x), rtol=1e-5, atol=1e-6)
        gru.convz1.weight.mul_(0.5)
        plain.convz1.weight.mul_(0.5)
        torch.testing.assert_close(gru(h, x), plain(h, x), rtol=1e-5, atol=1e-6)
    gru.train()
    gru(h, x).sum().backward()
    assert gru.convz1.weight.grad is not None and gru.convr2.weight.grad is not None


def test_graph_key_changes_with_weights_and_switches():
    """everything a captured graph bakes in is part of its cache key."""
    import torch
    from nndepth_b200.raft_stereo import BaseRAFTStereo
    model = BaseRAFTStereo(iters=2).eval()
    x = torch.zeros(1, 3, 64, 64)
    k0 = model._graph_key(x)
    assert model._graph_key(x) == k0
    model._graphs["sentinel"] = 1
    model.load_state_dict(model.state_dict())
    assert model._graphs == {} and model._graph_key(x) != k0          # versions bumped, cache dropped
    k1 = model._graph_key(x)
    for attr, value in (("fuse_gru", False), ("fuse_motion_front", False), ("fp16_encoder", False), ("final_only", True),
                        ("dense_precision", "mixed16")):
        old = getattr(model, attr, None)
        setattr(model, attr, value)
        assert model._graph_key(x) != k1, attr
        setattr(model, attr, old)
    model._graphs["sentinel"] = 1
    model.float()
    assert model._graphs == {}

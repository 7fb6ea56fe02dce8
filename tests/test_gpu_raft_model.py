"""GPU parity of the full RAFT-Stereo forward on the B200 kernels vs the reference model's outputs.

Bar (BASELINE.json): final disparity end-point error within 0.01 px of the reference.
"""
import numpy as np
import pytest
import torch

from helpers import epe, seeded_pair, state_fingerprint

pytestmark = pytest.mark.gpu
EPE_BAR = 0.01


def build(golden_case, final_only=False):
    from nndepth_b200.raft_stereo import BaseRAFTStereo
    torch.manual_seed(0)
    model = BaseRAFTStereo(iters=int(golden_case["iters"])).eval()
    np.testing.assert_allclose(state_fingerprint(model), golden_case["fingerprint"], rtol=1e-12)
    model = model.cuda()
    model.final_only = final_only
    return model


@pytest.mark.parametrize("tf32_convs", [False, True])
def test_small_all_iterations(golden, tf32_convs):
    g = golden("raft_small")
    model = build(g)
    left, right = (t.cuda() for t in seeded_pair(g["shape"]))
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = tf32_convs
    try:
        with torch.no_grad():
            outs = model(left, right)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    ref = torch.from_numpy(g["all_up_disp"]).cuda()
    assert len(outs) == int(g["iters"])
    for i, o in enumerate(outs):
        assert o["up_disp"].shape == ref[i].shape
        assert epe(o["up_disp"], ref[i]) < EPE_BAR, (i, epe(o["up_disp"], ref[i]))


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_kitti_32_iterations(golden, precision):
    """BASELINE config 2 geometry (384x1248 padded KITTI, 32 iterations), one pair."""
    import nndepth_b200 as nb
    g = golden("raft_kitti")
    model = build(g, final_only=True)
    model.update_block.gru.fuse_gates()
    left, right = (t.cuda() for t in seeded_pair(g["shape"]))
    old_prec = nb.get_volume_precision()
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    nb.set_volume_precision(precision)
    try:
        with torch.no_grad():
            out = model(left, right)[-1]["up_disp"]
            graphed = model.forward_graphed(left, right)[-1]["up_disp"].clone()
    finally:
        torch.backends.cudnn.allow_tf32 = old
        nb.set_volume_precision(old_prec)
    ref = torch.from_numpy(g["final_up_disp"]).cuda()
    assert out.shape == ref.shape == (1, 1, 384, 1248)
    assert epe(out, ref) < EPE_BAR, epe(out, ref)
    assert epe(graphed, ref) < EPE_BAR, epe(graphed, ref)


@pytest.mark.parametrize("mode", ["mixed16", "fp32"])
def test_final_only_gives_the_same_last_prediction(golden, mode):
    """``final_only`` skips the mask head and the upsampling of every iteration but the last: the mask
    never feeds the recurrence, so the last prediction must be BIT-identical to the full forward's."""
    g = golden("raft_kitti")
    left, right = (t.cuda() for t in seeded_pair(g["shape"]))
    outs = []
    for final_only in (False, True):
        model = build(g, final_only=final_only)
        model.dense_precision = mode
        with torch.no_grad():
            preds = model(left, right)
        assert len(preds) == (1 if final_only else model.iters)
        outs.append(preds[-1]["up_disp"])
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("mode,bar", [("fp32", 0.001), ("mixed", EPE_BAR), ("mixed2x", EPE_BAR), ("mixed16", EPE_BAR)])
def test_kitti_dense_precision_modes(golden, mode, bar):
    """The bench's dense-layer precision modes against the reference disparity (KITTI geometry, 32 iterations):
    "mixed" (ConvGRU fp32, other convolutions TF32) must stay inside the 0.01 px bar, "fp32" far inside."""
    g = golden("raft_kitti")
    model = build(g, final_only=True)
    model.dense_precision = mode
    left, right = (t.cuda() for t in seeded_pair(g["shape"]))
    before = torch.backends.cudnn.allow_tf32
    with torch.no_grad():
        out = model(left, right)[-1]["up_disp"]
        graphed = model.forward_graphed(left, right)[-1]["up_disp"].clone()
    assert torch.backends.cudnn.allow_tf32 == before          # the mode does not leak into global state
    ref = torch.from_numpy(g["final_up_disp"]).cuda()
    assert epe(out, ref) < bar, epe(out, ref)
    assert epe(graphed, ref) < bar, epe(graphed, ref)


def test_fused_gru_matches_unfused_weight_split():
    """The fused channels-last ConvGRU runner vs the same weight-split TF32 recurrence written with torch ops, and vs fp32."""
    from nndepth_b200.raft_stereo import SepConvGRU, cudnn_tf32
    torch.manual_seed(7)
    gru = SepConvGRU(hidden_dim=128, input_dim=256).cuda().eval()
    N, H, W = 2, 12, 20
    h0 = torch.tanh(torch.randn(N, 128, H, W, device="cuda"))
    inp = torch.relu(torch.randn(N, 128, H, W, device="cuda"))
    motions = [torch.randn(N, 128, H, W, device="cuda") for _ in range(3)]
    with torch.no_grad():
        gru.recurrence = "wsplit"
        run = gru.start(h0, inp)
        h_ref, h_fp32 = h0, h0
        for m in motions:
            fused = run.step(m)
            h_ref = gru(h_ref, torch.cat([inp, m], 1))
            gru.recurrence = "fp32"
            h_fp32 = gru(h_fp32, torch.cat([inp, m], 1))
            gru.recurrence = "wsplit"
            assert fused.shape == h_ref.shape
            assert (fused - h_ref).abs().max().item() < 2e-5        # same arithmetic, fused vs torch ops
            assert (fused - h_fp32).abs().max().item() < 1e-3       # activations are rounded to TF32 (2^-11 relative)


def test_fused_gru_fp16_matches_torch_ops_and_fp32():
    """The fp16 form of the fused runner (fp16 staging / split weights / pre-activations, fp32 gates) vs the same
    recurrence in torch ops, and vs fp32: fp16 carries TF32's mantissa, so the distance to fp32 is the same."""
    from nndepth_b200.raft_stereo import SepConvGRU
    torch.manual_seed(7)
    gru = SepConvGRU(hidden_dim=128, input_dim=256).cuda().eval()
    N, H, W = 2, 12, 20
    h0 = torch.tanh(torch.randn(N, 128, H, W, device="cuda"))
    inp = torch.relu(torch.randn(N, 128, H, W, device="cuda"))
    motions = [torch.randn(N, 128, H, W, device="cuda") for _ in range(3)]
    with torch.no_grad():
        gru.recurrence = "wsplit16"
        run = gru.start(h0, inp)
        assert run.half and run.S.dtype == torch.float16
        h_ref, h_fp32 = h0, h0
        for i, m in enumerate(motions):
            # channels-last fp32, NCHW fp32 and channels-last fp16 sources all stage to the same rows
            src = [m.contiguous(memory_format=torch.channels_last), m, m.half().contiguous(memory_format=torch.channels_last)][i]
            fused = run.step(src)
            h_ref = gru(h_ref, torch.cat([inp, m], 1))
            gru.recurrence = "fp32"
            h_fp32 = gru(h_fp32, torch.cat([inp, m], 1))
            gru.recurrence = "wsplit16"
            assert fused.dtype == torch.float32 and fused.shape == h_ref.shape
            assert (fused - h_ref).abs().max().item() < 2e-3        # fp16 pre-activations: summation order moves a rounding
            assert (fused - h_fp32).abs().max().item() < 2e-3       # 2^-11 relative on operands and pre-activations


def test_engine_bench_configuration_stays_inside_the_bar(golden):
    """The exact configuration bench.py times -- StereoEngine (channels-last encoder, CUDA graph, fused GRU glue,
    fused lookup / upsampling), dense_precision = "mixed16" -- against the reference disparity."""
    from nndepth_b200.engine import StereoEngine
    from nndepth_b200.raft_stereo import BaseRAFTStereo
    g = golden("raft_kitti")
    torch.manual_seed(0)
    model = BaseRAFTStereo(iters=int(g["iters"])).eval()
    model.dense_precision = "mixed16"
    engine = StereoEngine(model, device="cuda", use_cuda_graph=True)
    left, right = (t.cuda() for t in seeded_pair(g["shape"]))
    ref = torch.from_numpy(g["final_up_disp"]).cuda()
    for _ in range(2):                                  # second call replays the captured graph
        out = engine.infer_device(left, right)
        assert out.shape == ref.shape
        assert epe(out, ref) < EPE_BAR, epe(out, ref)
    host = engine.infer(left.cpu().pin_memory(), right.cpu().pin_memory())
    assert epe(host.cuda(), ref) < EPE_BAR
    # pipelined host-to-host path (two batches in flight on a copy stream): same numbers as the synchronous call, and
    # a different batch in the other slot does not disturb them
    hl, hr = left.cpu().pin_memory(), right.cpu().pin_memory()
    t1 = engine.submit(hl, hr)
    t2 = engine.submit(hr, hl)
    first = engine.collect(t1).clone()
    t3 = engine.submit(hl, hr)
    engine.collect(t2)
    third = engine.collect(t3).clone()
    assert torch.equal(first, host) and torch.equal(third, host)


@pytest.mark.parametrize("mode", ["mixed2x", "mixed16"])
@pytest.mark.parametrize("graph", [False, True])
def test_small_all_iterations_bench_mode(golden, graph, mode):
    """The bench's mode (mixed2x, every fusion on) at the small golden shape: every iteration's upsampled map
    against the reference, eager and as a CUDA graph (few pixel groups, ragged tiles, 12 iterations)."""
    from nndepth_b200.engine import StereoEngine
    from nndepth_b200.raft_stereo import BaseRAFTStereo
    g = golden("raft_small")
    torch.manual_seed(0)
    model = BaseRAFTStereo(iters=int(g["iters"])).eval()
    model.dense_precision = mode
    engine = StereoEngine(model, device="cuda", use_cuda_graph=graph)
    left, right = (t.cuda() for t in seeded_pair(g["shape"]))
    ref = torch.from_numpy(g["all_up_disp"]).cuda()
    with torch.no_grad():
        outs = engine.model.forward_graphed(left, right) if graph else engine.model(left, right)
    assert len(outs) == int(g["iters"])
    for i, o in enumerate(outs):
        assert o["up_disp"].shape == ref[i].shape
        assert epe(o["up_disp"], ref[i]) < EPE_BAR, (i, epe(o["up_disp"], ref[i]))


def test_graph_replay_after_weight_reload_uses_the_new_weights():
    """derived weights (rounded / folded / split copies) are captured by address; after load_state_dict a
    replay must not mix fresh parameters with stale derived tensors."""
    from nndepth_b200.raft_stereo import BaseRAFTStereo
    torch.manual_seed(0)
    model = BaseRAFTStereo(iters=4).eval().cuda()
    model.dense_precision = "mixed16"
    model.update_block.gru.fuse_gates()
    torch.manual_seed(11)
    other = BaseRAFTStereo(iters=4).eval().cuda()
    other.dense_precision = "mixed16"
    left, right = (t.cuda() for t in seeded_pair((1, 3, 128, 256)))
    with torch.no_grad():
        first = model.forward_graphed(left, right)[-1]["up_disp"].clone()
        model.forward_graphed(left, right)                          # a replay
        model.load_state_dict(other.state_dict())
        reloaded = model.forward_graphed(left, right)[-1]["up_disp"].clone()
        want = other(left, right)[-1]["up_disp"]
    assert not torch.equal(first, reloaded)
    assert epe(reloaded, want) < 1e-4, epe(reloaded, want)


def test_dense_precision_does_not_leak_into_later_calls():
    """after a mixed16 run, ``dense_precision = None`` runs on the caller's own flags again (fp32 recurrence,
    fp32 encoder), eagerly and in a freshly captured graph."""
    from nndepth_b200.raft_stereo import BaseRAFTStereo
    torch.manual_seed(0)
    model = BaseRAFTStereo(iters=4).eval().cuda()
    left, right = (t.cuda() for t in seeded_pair((1, 3, 128, 256)))
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            before = model(left, right)[-1]["up_disp"].clone()
            model.dense_precision = "mixed16"
            model(left, right)
            assert model.update_block.gru.recurrence is None and not getattr(model.fnet, "half_convs", False)
            model.dense_precision = None
            after = model(left, right)[-1]["up_disp"].clone()
            graphed = model.forward_graphed(left, right)[-1]["up_disp"].clone()
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert torch.equal(before, after)
    assert epe(graphed, before) < 1e-5

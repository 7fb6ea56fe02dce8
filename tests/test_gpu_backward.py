"""GPU parity of the training path: gradients of CorrBlock1D w.r.t. the feature maps vs torch autograd through a
plain-torch restatement of the reference chain (matmul / avg_pool1d / gather lerp) on the device."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def torch_chain(f1, f2, coords_list, L=4, r=4):
    """reference raft_stereo/cost_volume.py:12-61 + utils.py:4-27 in differentiable torch ops (double precision)."""
    B, C, H, W = f1.shape
    vol = torch.matmul(f1.permute(0, 2, 3, 1), f2.permute(0, 2, 1, 3)) / C ** 0.5
    lvl = vol.reshape(B * H * W, 1, -1)
    pyr = [lvl]
    for _ in range(L):
        lvl = F.avg_pool1d(lvl, 2)
        pyr.append(lvl)
    outs = []
    dx = torch.linspace(-r, r, 2 * r + 1, device=f1.device, dtype=f1.dtype).view(1, -1)
    for coords in coords_list:
        per = []
        for l in range(L):
            rows = pyr[l].reshape(B * H * W, -1)
            x = dx + coords.reshape(-1, 1).to(f1.dtype) / 2 ** l
            w2 = rows.shape[1]
            t = torch.clamp(x / (w2 - 1), 0, 1) * (w2 - 1)
            i0, i1 = t.floor().long(), t.ceil().long()
            coef = i1 - t
            per.append((coef * rows.gather(1, i0) + (1 - coef) * rows.gather(1, i1)).view(B, H, W, -1))
        outs.append(torch.cat(per, -1).permute(0, 3, 1, 2))
    return outs


@pytest.mark.parametrize("shape", [(2, 16, 3, 40), (1, 32, 2, 156)])
@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_gradients_match_torch_autograd(shape, precision):
    import nndepth_b200 as nb
    B, C, H, W = shape
    torch.manual_seed(B * C + W)
    f1 = torch.randn(B, C, H, W, device="cuda", requires_grad=True)
    f2 = torch.randn(B, C, H, W, device="cuda", requires_grad=True)
    grid = torch.arange(W, device="cuda").float().view(1, 1, 1, W).repeat(B, 1, H, 1)
    coords_list = [grid - torch.rand(B, 1, H, W, device="cuda") * 12, grid + 3.3, grid - 100.0]
    weights = [torch.randn(B, 36, H, W, device="cuda") for _ in coords_list]

    blk = nb.CorrBlock1D(f1, f2, 4, 4, precision=precision)
    loss = sum((blk(c) * w).sum() for c, w in zip(coords_list, weights))
    loss.backward()
    g1, g2 = f1.grad.clone(), f2.grad.clone()

    d1 = f1.detach().double().requires_grad_(True)
    d2 = f2.detach().double().requires_grad_(True)
    ref_loss = sum((o * w.double()).sum() for o, w in zip(torch_chain(d1, d2, coords_list), weights))
    ref_loss.backward()
    tol = 1e-5 if precision == "fp32" else 2e-3
    assert abs(loss.item() - ref_loss.item()) <= tol * max(1.0, abs(ref_loss.item())) * 10
    for g, ref in ((g1, d1.grad), (g2, d2.grad)):
        scale = ref.abs().max().item()
        assert (g.double() - ref).abs().max().item() <= 1e-4 * scale


def test_inference_path_is_unchanged_and_grad_free():
    import nndepth_b200 as nb
    f = torch.randn(1, 8, 2, 16, device="cuda", requires_grad=True)
    with torch.no_grad():
        blk = nb.CorrBlock1D(f, f, 2, 4)
        out = blk(torch.zeros(1, 1, 2, 16, device="cuda"))
        grp = nb.GroupCorrBlock1D(f, f, 2, 4, 2)
    assert not out.requires_grad and blk._graph_buffer is None
    assert grp._graph_buffer is None


# ------------------------------------------------------------------------------------------------------------------
# The other differentiable blocks against autograd through the UNMODIFIED reference classes in float64 (oracle/_ref)
# ------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ref():
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("oracle/_ref is not staged: run __graft_entry__.build() in the build container first")
    ref_shim.install()
    return ref_shim


def _close(got, want, tol=1e-4):
    scale = want.abs().max().item()
    err = (got.double() - want).abs().max().item()
    assert err <= tol * max(scale, 1e-12), (err, scale)


def test_group_corr_gradients(ref):
    import nndepth_b200 as nb
    from nndepth.models.raft_stereo.cost_volume import GroupCorrBlock1D as RefGroup
    B, C, H, W, G = 2, 32, 3, 36, 4
    torch.manual_seed(5)
    f1 = torch.randn(B, C, H, W, device="cuda", requires_grad=True)
    f2 = torch.randn(B, C, H, W, device="cuda", requires_grad=True)
    grid = torch.arange(W, device="cuda").float().view(1, 1, 1, W).repeat(B, 1, H, 1)
    coords_list = [grid - torch.rand(B, 1, H, W, device="cuda") * 9, grid - 60.0]
    weights = [torch.randn(B, 4 * G * 9, H, W, device="cuda") for _ in coords_list]
    blk = nb.GroupCorrBlock1D(f1, f2, 4, 4, G)
    loss = sum((blk(c) * w).sum() for c, w in zip(coords_list, weights))
    loss.backward()
    d1, d2 = f1.detach().double().requires_grad_(True), f2.detach().double().requires_grad_(True)
    rblk = RefGroup(d1, d2, 4, 4, G)
    rloss = sum((rblk(c.double()).double() * w.double()).sum() for c, w in zip(coords_list, weights))
    rloss.backward()
    assert abs(loss.item() - rloss.item()) <= 1e-4 * max(1.0, abs(rloss.item()))
    _close(f1.grad, d1.grad)
    _close(f2.grad, d2.grad)
    assert f1.grad[:, G * G:].abs().max().item() == 0.0        # only the first G*G channels take part (split quirk)


def test_igev_volume_gradients(ref):
    """GeometryAwareCostVolume under grad: gradients for the feature maps and the 3-D regulariser's parameters, through
    the dual lookup AND through geo_aware_cv[0] (what the model's cv_squeezer reads, igev_stereo/model.py:144)."""
    import copy
    import nndepth_b200 as nb
    from nndepth.models.igev_stereo.cost_volume import GeometryAwareCostVolume as RefGEV

    class Reg(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = torch.nn.Conv3d(8, 8, 3, padding=1)

        def forward(self, vol, feats):
            return torch.tanh(self.conv(vol)) + 0.1 * vol

    B, C, H, W, G = 1, 64, 2, 24, 8
    torch.manual_seed(9)
    reg = Reg().cuda()
    reg64 = copy.deepcopy(reg).double()
    f1 = torch.randn(B, C, H, W, device="cuda", requires_grad=True)
    f2 = torch.randn(B, C, H, W, device="cuda", requires_grad=True)
    grid = torch.arange(W, device="cuda").float().view(1, 1, 1, W).repeat(B, 1, H, 1)
    coords = grid - torch.rand(B, 1, H, W, device="cuda") * 6
    w_out = torch.randn(B, 576, H, W, device="cuda")
    w_geo = torch.randn(B * G * H * W, 1, W, device="cuda")

    old_flags = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False     # the regulariser is cuDNN: keep it fp32
    try:
        cv = nb.GeometryAwareCostVolume(f1, f2, [], reg, 4, 4, G)
        loss = (cv(coords) * w_out).sum() + (cv.geo_aware_cv[0] * w_geo).sum()
        loss.backward()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old_flags

    d1, d2 = f1.detach().double().requires_grad_(True), f2.detach().double().requires_grad_(True)
    rcv = RefGEV(d1, d2, [], reg64, 4, 4, G)
    rloss = (rcv(coords.double()).double() * w_out.double()).sum() + (rcv.geo_aware_cv[0] * w_geo.double()).sum()
    rloss.backward()
    assert abs(loss.item() - rloss.item()) <= 1e-4 * max(1.0, abs(rloss.item()))
    _close(f1.grad, d1.grad)
    _close(f2.grad, d2.grad)
    _close(reg.conv.weight.grad, reg64.conv.weight.grad)
    _close(reg.conv.bias.grad, reg64.conv.bias.grad)


@pytest.mark.parametrize("small_patch", [False, True])
@pytest.mark.parametrize("iter_mode", [False, True, "through_warp"])
def test_agcl_gradients(ref, small_patch, iter_mode, monkeypatch):
    """AGCL under grad (both modes, both windows): gradients for both feature maps, the flow and (offset mode) the learned
    offsets vs autograd through the reference's AGCL / bilinear_grid_sample in float64.  In iter mode the reference pads
    a DETACHED clone of the warped right map (manual_pad, cre_stereo/utils.py:29-31): only the left features receive a
    gradient, and so it is here; "through_warp" checks the full backward against the reference with that detach removed."""
    import nndepth_b200 as nb
    import nndepth.models.cre_stereo.cost_volume as ref_cv
    from nndepth.models.cre_stereo.cost_volume import AGCL as RefAGCL
    through_warp = iter_mode == "through_warp"
    iter_mode = bool(iter_mode)
    if through_warp:
        monkeypatch.setattr(ref_cv, "manual_pad", lambda x, pady, padx: F.pad(x, (padx, padx, pady, pady), "replicate"))
    N, C, H, W = 2, 32, 7, 11
    torch.manual_seed(3 + int(small_patch) + 2 * int(iter_mode))
    f1 = torch.randn(N, C, H, W, device="cuda", requires_grad=True)
    f2 = torch.randn(N, C, H, W, device="cuda", requires_grad=True)
    flow = (torch.randn(N, 2, H, W, device="cuda") * 2).requires_grad_(True)
    extra = (torch.rand(N, 18, H, W, device="cuda") * 2 - 1).requires_grad_(True)
    w_out = torch.randn(N, 36, H, W, device="cuda")
    blk = nb.AGCL(f1, f2)
    blk.detach_warped = not through_warp
    out = blk(flow, None if iter_mode else extra, small_patch, iter_mode)
    loss = (out * w_out).sum()
    loss.backward()
    dd = [t.detach().double().requires_grad_(True) for t in (f1, f2, flow, extra)]
    rout = RefAGCL(dd[0], dd[1])(dd[2], None if iter_mode else dd[3], small_patch, iter_mode)
    rloss = (rout.double() * w_out.double()).sum()
    rloss.backward()
    assert abs(loss.item() - rloss.item()) <= 1e-4 * max(1.0, abs(rloss.item()))
    _close(f1.grad, dd[0].grad)
    if iter_mode and not through_warp:
        assert dd[1].grad is None and dd[2].grad is None          # the reference's detach
        assert f2.grad is None and flow.grad is None
        return
    _close(f2.grad, dd[1].grad)
    _close(flow.grad, dd[2].grad, tol=1e-3)
    if not iter_mode:
        _close(extra.grad, dd[3].grad, tol=1e-3)


def test_agcl_gradients_reach_the_attention_module(ref):
    """The cascade's coarsest scale passes a LoFTR cross-attention module (cre_stereo/model.py:200): its parameters train."""
    import nndepth_b200 as nb

    class Att(torch.nn.Module):
        def __init__(self, C):
            super().__init__()
            self.a, self.b = torch.nn.Linear(C, C), torch.nn.Linear(C, C)

        def forward(self, x, y):
            return x + torch.tanh(self.a(y)), y + torch.tanh(self.b(x))

    N, C, H, W = 1, 32, 5, 8
    torch.manual_seed(1)
    att = Att(C).cuda()
    f1, f2 = torch.randn(N, C, H, W, device="cuda"), torch.randn(N, C, H, W, device="cuda")
    out = nb.AGCL(f1, f2, att=att)(torch.zeros(N, 2, H, W, device="cuda"), torch.zeros(N, 18, H, W, device="cuda"), False, False)
    out.sum().backward()
    assert att.a.weight.grad is not None and att.a.weight.grad.abs().sum().item() > 0
    assert att.b.weight.grad is not None and att.b.weight.grad.abs().sum().item() > 0

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
    assert not out.requires_grad and blk._graph_buffer is None
    # the grouped / IGEV / AGCL blocks stay inference-only and say so
    with pytest.raises(RuntimeError, match="inference-only"):
        nb.GroupCorrBlock1D(f, f, 2, 4, 2)

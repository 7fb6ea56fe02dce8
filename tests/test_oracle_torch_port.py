"""Pin oracle/torch_port.py (the multi-threaded CPU baseline) to the reference goldens, and check the
RAFT-Stereo model shell reproduces the reference's full-model outputs on CPU with it."""
import numpy as np
import pytest
import torch

from oracle import torch_port as tp
from helpers import epe, seeded_pair, state_fingerprint

CASES = ["corr1d_small", "corr1d_odd", "corr1d_kitti_row", "corr1d_r3l3"]


@pytest.mark.parametrize("case", CASES)
def test_torch_port_matches_reference(golden, case):
    g = golden(case)
    L, r = int(g["num_levels"]), int(g["radius"])
    torch.set_num_threads(1)
    blk = tp.CorrBlock1D(torch.from_numpy(g["fmap1"]), torch.from_numpy(g["fmap2"]), L, r)
    scale = np.abs(g["pyr0"]).max()
    for l in range(L + 1):
        np.testing.assert_allclose(blk.corr_pyramid[l].numpy(), g[f"pyr{l}"], rtol=1e-5, atol=1e-5 * scale)
    blk.corr_pyramid = [torch.from_numpy(g[f"pyr{l}"]) for l in range(L + 1)]
    for regime in ("int", "sub", "oob"):
        out = blk(torch.from_numpy(g[f"coords_{regime}"])).numpy()
        np.testing.assert_array_equal(out, g[f"out_{regime}"])


def test_model_shell_draws_reference_weights(golden):
    from nndepth_b200.raft_stereo import BaseRAFTStereo
    g = golden("raft_small")
    torch.manual_seed(0)
    model = BaseRAFTStereo(iters=int(g["iters"])).eval()
    np.testing.assert_allclose(state_fingerprint(model), g["fingerprint"], rtol=1e-12)


def test_model_shell_matches_reference_outputs_on_cpu(golden):
    """Shell (plain torch layers) + CPU oracle correlation == the reference model's outputs."""
    from nndepth_b200.raft_stereo import BaseRAFTStereo
    g = golden("raft_small")
    torch.manual_seed(0)
    model = BaseRAFTStereo(iters=int(g["iters"])).eval()
    model.corr_fn = tp.CorrBlock1D
    left, right = seeded_pair(g["shape"])
    with torch.no_grad():
        outs = model(left, right)
    assert isinstance(outs, list) and len(outs) == int(g["iters"])
    ref = torch.from_numpy(g["all_up_disp"])
    for i, o in enumerate(outs):
        assert o["up_disp"].shape == ref[i].shape
        assert epe(o["up_disp"], ref[i]) < 1e-4, i
    # gate fusion and final-only mode do not change the result
    model.update_block.gru.fuse_gates()
    model.final_only = True
    with torch.no_grad():
        last = model(left, right)
    assert len(last) == 1 and epe(last[0]["up_disp"], ref[-1]) < 1e-4

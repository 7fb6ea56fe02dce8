"""Shared helpers of the test-suite (not collected as tests)."""
import numpy as np
import torch


def state_fingerprint(model):
    """Same formula as tests/golden/make_goldens.py:state_fingerprint."""
    acc_abs, acc_ramp, count = 0.0, 0.0, 0
    for _, t in model.state_dict().items():
        t = t.detach().double().cpu().reshape(-1)
        acc_abs += float(t.abs().sum())
        acc_ramp += float((t * torch.linspace(-1, 1, t.numel(), dtype=torch.float64)).sum())
        count += t.numel()
    return np.float64([acc_abs, acc_ramp, count])


def seeded_pair(shape, seed=1):
    """The synthetic stereo pair of make_goldens.py:raft_model_cases (CPU generator, U(-1,1))."""
    gen = torch.Generator().manual_seed(seed)
    left = torch.rand(tuple(int(s) for s in shape), generator=gen) * 2 - 1
    right = torch.rand(tuple(int(s) for s in shape), generator=gen) * 2 - 1
    return left, right


def epe(a, b):
    """Mean end-point error, reference raft_stereo/loss.py:40-43 (1 channel: |a - b|)."""
    return (a - b).abs().mean().item()

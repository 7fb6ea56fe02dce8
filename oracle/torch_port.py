"""Oracle (TEST / BASELINE INFRASTRUCTURE ONLY): multi-threaded CPU restatement of the correlation path.

Same algorithm as ``oracle/corr1d.py`` (numpy, single-threaded), restated with torch CPU tensor ops so
that it can use every host core -- the reference itself is a chain of ATen CPU ops, so this is the
fair "reference CPU implementation" for ``bench.py --impl reference`` and the ``cpu_baseline`` leg.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` may import it; the product package
``nndepth_b200`` never does.

Follows ``nndepth/models/raft_stereo/cost_volume.py:7-61`` and ``raft_stereo/utils.py:4-27`` of the
reference; pinned bit-for-bit against ``tests/golden/corr1d_*.npz`` by ``tests/test_oracle_torch_port.py``.
"""
import torch


def all_pairs_correlation(fmap1, fmap2):
    """``(B,H,W1,W2)``: per-row ``f1^T f2`` then a true division by ``C**0.5`` (cost_volume.py:55-61)."""
    channels = fmap1.shape[1]
    rows1 = fmap1.float().permute(0, 2, 3, 1)      # (B, H, W1, C)
    rows2 = fmap2.float().permute(0, 2, 1, 3)      # (B, H, C, W2)
    return torch.matmul(rows1, rows2) / channels ** 0.5


def avg_pool_pairs(level):
    """``avg_pool1d(x, 2)`` on the last axis of ``(rows, w)``: ``(x[2j] + x[2j+1]) / 2``, odd tail dropped."""
    half = level.shape[-1] // 2
    return (level[:, 0:2 * half:2] + level[:, 1:2 * half:2]) / 2


def linear_sampler(rows, x):
    """``rows (N, w2)`` sampled at ``x (N, T)``; clamp-normalised positions (utils.py:15-27)."""
    span = rows.shape[1] - 1
    t = torch.clamp(x / span, 0, 1) * span
    lo, hi = t.floor(), t.ceil()
    coef = hi - t
    return coef * rows.gather(1, lo.long()) + (1 - coef) * rows.gather(1, hi.long())


class CorrBlock1D:
    """CPU twin of the reference class: list-of-levels pyramid, per-level lookup, NCHW output."""

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4):
        self.num_levels, self.radius = num_levels, radius
        volume = all_pairs_correlation(fmap1, fmap2)
        level = volume.reshape(-1, volume.shape[-1])
        self.corr_pyramid = [level]
        for _ in range(num_levels):
            level = avg_pool_pairs(level)
            self.corr_pyramid.append(level)

    def __call__(self, coords):
        B, _, H, W = coords.shape
        taps = torch.linspace(-self.radius, self.radius, 2 * self.radius + 1).view(1, -1)
        centre = coords.reshape(-1, 1)
        sampled = [linear_sampler(self.corr_pyramid[lvl], taps + centre / 2 ** lvl).view(B, H, W, -1)
                   for lvl in range(self.num_levels)]
        return torch.cat(sampled, dim=-1).permute(0, 3, 1, 2).contiguous().float()

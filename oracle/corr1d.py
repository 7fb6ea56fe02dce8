"""Oracle (test infrastructure, numpy): RAFT-Stereo 1-D correlation pyramid and its lookup.

Restates ``nndepth/models/raft_stereo/cost_volume.py`` and ``nndepth/models/raft_stereo/utils.py``
(identical copy: ``nndepth/models/igev_stereo/utils.py``) of the reference.  Parity is pinned by
``tests/golden/corr1d_*.npz`` (outputs of the unmodified reference; see ``tests/golden/make_goldens.py``).

Shapes follow the reference: features are NCHW ``(B, C, H, W)``; the volume is ``(B, H, W1, W2)``;
pyramid level ``l`` is kept as a 2-D array ``(B*H*W1, w_l)``.
"""
import numpy as np

F32 = np.float32


def all_pairs_correlation(fmap1, fmap2, accumulate="f64"):
    """``corr[b,h,w1,w2] = sum_c f1[b,c,h,w1] * f2[b,c,h,w2] / C**0.5``.

    Reference: ``CorrBlock1D.corr`` raft_stereo/cost_volume.py:55-61 (permute + ``torch.matmul``
    then a true division by the Python float ``C**0.5``).  The contraction is accumulated in fp64
    and rounded once to fp32 (``accumulate="f64"``, the default: the reference's fp32 SGEMM lies
    within ``C * 2**-24`` relative of it, whichever order it sums in) or in fp32 (``"f32"``).  The
    scale is applied as an fp32 division, as in the reference.
    """
    f1 = np.asarray(fmap1, dtype=F32)
    f2 = np.asarray(fmap2, dtype=F32)
    assert f1.ndim == 4 and f2.ndim == 4 and f1.shape[:3] == f2.shape[:3]
    C = f1.shape[1]
    acc_t = np.float64 if accumulate == "f64" else F32
    dot = np.einsum("bchi,bchj->bhij", f1.astype(acc_t), f2.astype(acc_t), optimize=True)
    return dot.astype(F32) / F32(C ** 0.5)


def avg_pool_pairs(level):
    """``F.avg_pool1d(x, 2)`` over the last axis: kernel 2, stride 2, odd tail dropped.

    Reference: raft_stereo/cost_volume.py:32-34 (and igev_stereo/cost_volume.py:48-52).  ATen sums
    the two taps and divides by the window size, i.e. ``(a + b) / 2`` in fp32.
    """
    level = np.asarray(level, dtype=F32)
    half = level.shape[-1] // 2
    return (level[..., 0:2 * half:2] + level[..., 1:2 * half:2]) / F32(2)


def build_pyramid(volume, num_levels=4):
    """``num_levels + 1`` arrays ``(prod(leading dims), w_l)``; level 0 is the volume itself.

    Reference: ``CorrBlock1D.__init__`` raft_stereo/cost_volume.py:28-34 -- the reference keeps one
    more level than ``__call__`` ever reads (it loops ``range(num_levels)`` at :41).
    """
    volume = np.asarray(volume, dtype=F32)
    level = volume.reshape(-1, volume.shape[-1])
    pyramid = [level]
    for _ in range(num_levels):
        level = avg_pool_pairs(level)
        pyramid.append(level)
    return pyramid


def sampler_indices(x, w2):
    """Clamped sample position and its two integer neighbours.

    Reference: ``linear_sampler`` raft_stereo/utils.py:15-21::

        t = clamp(x / (w2 - 1), 0, 1) * (w2 - 1);  i0 = floor(t);  i1 = ceil(t)

    The fp32 divide-then-multiply does NOT round-trip integers, so both
    steps are kept as separate IEEE fp32 operations; i0/i1 are the bit-exact contract.
    """
    x = np.asarray(x, dtype=F32)
    if w2 < 2:
        raise ValueError("linear_sampler divides by (w2 - 1): level width must be >= 2")
    span = F32(w2 - 1)
    t = np.clip(x / span, F32(0), F32(1)) * span
    i0 = np.floor(t).astype(np.int64)
    i1 = np.ceil(t).astype(np.int64)
    return t, i0, i1


def linear_sampler(rows, x):
    """Linear interpolation of each row of ``rows (N, w2)`` at positions ``x (N, T)`` -> ``(N, T)``.

    Reference: ``linear_sampler`` raft_stereo/utils.py:4-27.  Border handling is a clamp of the
    coordinate (replicate), not zero padding.  ``coef = i1 - t``; result
    ``coef * v[i0] + (1 - coef) * v[i1]`` with every product and sum rounded to fp32 (no FMA).
    """
    rows = np.asarray(rows, dtype=F32)
    t, i0, i1 = sampler_indices(x, rows.shape[1])
    v0 = np.take_along_axis(rows, i0, axis=1)
    v1 = np.take_along_axis(rows, i1, axis=1)
    coef = i1.astype(F32) - t
    return coef * v0 + (F32(1) - coef) * v1


def level_positions(coords, level, radius):
    """``x = dx + coords / 2**level`` for ``dx = -r..r``: ``(N,)`` -> ``(N, 2r+1)`` (cost_volume.py:44-46)."""
    dx = np.linspace(-radius, radius, 2 * radius + 1, dtype=F32)[None, :]
    centre = np.asarray(coords, dtype=F32).reshape(-1, 1) / F32(2 ** level)
    return dx + centre


def lookup(pyramid, coords, num_levels=4, radius=4):
    """Pyramid lookup: coords ``(B,1,H,W)`` -> ``(B, num_levels*(2r+1), H, W)``, channel ``l*(2r+1)+k``.

    Reference: ``CorrBlock1D.__call__`` raft_stereo/cost_volume.py:36-53.
    """
    coords = np.asarray(coords, dtype=F32)
    B, one, H, W = coords.shape
    assert one == 1
    per_level = []
    for lvl in range(num_levels):
        rows = pyramid[lvl].reshape(B * H * W, -1)
        x = level_positions(coords, lvl, radius)
        per_level.append(linear_sampler(rows, x).reshape(B, H, W, -1))
    out = np.concatenate(per_level, axis=-1)
    return np.ascontiguousarray(out.transpose(0, 3, 1, 2), dtype=F32)


def lookup_indices(widths, coords, num_levels=4, radius=4):
    """Per-level ``(i0, i1)`` int64 arrays ``(B*H*W, 2r+1)`` -- the integer half of the lookup."""
    out = []
    for lvl in range(num_levels):
        _, i0, i1 = sampler_indices(level_positions(coords, lvl, radius), widths[lvl])
        out.append((i0, i1))
    return out


class CorrBlock1D:
    """Oracle twin of the reference class (raft_stereo/cost_volume.py:7-61)."""

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4, accumulate="f64"):
        self.num_levels = num_levels
        self.radius = radius
        self.corr_pyramid = build_pyramid(all_pairs_correlation(fmap1, fmap2, accumulate), num_levels)

    def __call__(self, coords):
        return lookup(self.corr_pyramid, coords, self.num_levels, self.radius)


# ---------------------------------------------------------------------------------------------
# GroupCorrBlock1D
# ---------------------------------------------------------------------------------------------

def group_all_pairs_correlation(fmap1, fmap2, num_groups=4, accumulate="f64"):
    """``(B, G, H, W1, W2)`` grouped volume with the reference's quirks.

    Reference: ``GroupCorrBlock1D.corr`` raft_stereo/cost_volume.py:113-128.
    ``torch.split(fmap, num_groups, dim=1)`` makes chunks *of size* ``num_groups`` and the loop reads
    only the first ``num_groups`` of them, so group g = channels ``[G*g, G*g + G)``; the scale is
    ``1 / C_total**0.5``.
    """
    f1 = np.asarray(fmap1, dtype=F32)
    f2 = np.asarray(fmap2, dtype=F32)
    C = f1.shape[1]
    G = num_groups
    acc_t = np.float64 if accumulate == "f64" else F32
    vols = []
    for g in range(G):
        a = f1[:, G * g:G * g + G].astype(acc_t)
        b = f2[:, G * g:G * g + G].astype(acc_t)
        vols.append(np.einsum("bchi,bchj->bhij", a, b, optimize=True).astype(F32) / F32(C ** 0.5))
    return np.stack(vols, axis=1)


def group_lookup(pyramid, coords, num_levels=4, radius=4, num_groups=4):
    """Reference: ``GroupCorrBlock1D.__call__`` raft_stereo/cost_volume.py:92-111.

    Rows are ordered ``[b][g][h][w]``; the sampled ``(B*G*H*W, 2r+1)`` block is then *viewed* as
    ``(B, H, W, G*(2r+1))`` (:108) -- a reinterpretation of memory, not a transpose.  Reproduced
    as is.
    """
    coords = np.asarray(coords, dtype=F32)
    B, _, H, W = coords.shape
    G = num_groups
    rep = np.broadcast_to(coords.reshape(B, 1, H, W), (B, G, H, W)).reshape(-1)
    per_level = []
    for lvl in range(num_levels):
        rows = pyramid[lvl].reshape(B * G * H * W, -1)
        x = level_positions(rep, lvl, radius)
        per_level.append(linear_sampler(rows, x).reshape(B, H, W, -1))
    out = np.concatenate(per_level, axis=-1)
    return np.ascontiguousarray(out.transpose(0, 3, 1, 2), dtype=F32)


class GroupCorrBlock1D:
    """Oracle twin of raft_stereo/cost_volume.py:64-128."""

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4, num_groups=4, accumulate="f64"):
        self.num_levels, self.radius, self.num_groups = num_levels, radius, num_groups
        vol = group_all_pairs_correlation(fmap1, fmap2, num_groups, accumulate)
        self.corr_pyramid = build_pyramid(vol, num_levels)

    def __call__(self, coords):
        return group_lookup(self.corr_pyramid, coords, self.num_levels, self.radius, self.num_groups)

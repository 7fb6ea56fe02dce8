"""Oracle (test infrastructure, numpy): CREStereo adaptive group correlation layer (AGCL).

Restates ``nndepth/models/cre_stereo/cost_volume.py:6-154`` and the samplers of
``nndepth/models/cre_stereo/utils.py:5-107``.  The optional LoFTR cross-attention (``att``,
cost_volume.py:91-99) is dense attention and OUT OF SCOPE: it enters as a callable on
``(N, H*W, C)`` arrays.  Pinned by ``tests/golden/agcl_*.npz``.
"""
import numpy as np

from .corr1d import F32

NUM_GROUPS = 4          # cost_volume.py:69-70,101-102: channels are split into 4 groups
SEARCH_NUM = 9          # cost_volume.py:113


def coords_grid(batch, ht, wd):
    """``(N, 2, H, W)`` with x in channel 0 and y in channel 1.  Reference: cre_stereo/utils.py:23-26."""
    ys, xs = np.meshgrid(np.arange(ht, dtype=F32), np.arange(wd, dtype=F32), indexing="ij")
    return np.broadcast_to(np.stack([xs, ys])[None], (batch, 2, ht, wd)).copy()


def pixel_round_trip(p, size):
    """Pixel coordinate -> normalised [-1, 1] -> pixel coordinate, in fp32, as the reference does.

    Reference: ``bilinear_sampler`` utils.py:9-10 (``2 * p / (size - 1) - 1``) followed by
    ``bilinear_grid_sample`` utils.py:59-60 (``((g + 1) / 2) * (size - 1)``, align_corners=True).  The
    round trip is not the identity in fp32, and ``floor`` of the result picks the integer corner.
    """
    p = np.asarray(p, dtype=F32)
    span = F32(size - 1)
    g = (F32(2) * p) / span - F32(1)
    return ((g + F32(1)) / F32(2)) * span


def bilinear_sample_zero_pad(img, x, y):
    """Bilinear sample of ``img (N,C,H,W)`` at pixel positions ``x, y (N, P)`` -> ``(N, C, P)``.

    Reference: ``bilinear_grid_sample`` utils.py:34-107 == ``F.grid_sample(bilinear, zeros,
    align_corners=True)``: ``x0 = floor(x)``, ``x1 = x0 + 1``; weights from the unclamped corners
    (:71-74); the image is padded by one zero ring and corner indices are clamped into the padded
    range (:77-93), i.e. any corner outside the image contributes 0.
    Sum order ``Ia*wa + Ib*wb + Ic*wc + Id*wd`` (:107) with a=(x0,y0) b=(x0,y1) c=(x1,y0) d=(x1,y1).
    """
    img = np.asarray(img, dtype=F32)
    N, C, H, W = img.shape
    x0f = np.floor(x)
    y0f = np.floor(y)
    x0 = x0f.astype(np.int64)
    y0 = y0f.astype(np.int64)
    x1, y1 = x0 + 1, y0 + 1
    x1f, y1f = x1.astype(F32), y1.astype(F32)
    wa = (x1f - x) * (y1f - y)
    wb = (x1f - x) * (y - y0f)
    wc = (x - x0f) * (y1f - y)
    wd = (x - x0f) * (y - y0f)
    padded = np.zeros((N, C, H + 2, W + 2), dtype=F32)
    padded[:, :, 1:-1, 1:-1] = img
    flat = padded.reshape(N, C, -1)

    def corner(xi, yi):
        xi = np.clip(xi + 1, 0, W + 1)
        yi = np.clip(yi + 1, 0, H + 1)
        idx = (xi + yi * (W + 2))[:, None, :]
        return np.take_along_axis(flat, np.broadcast_to(idx, (N, C, idx.shape[-1])), axis=2)

    Ia, Ib, Ic, Id = corner(x0, y0), corner(x0, y1), corner(x1, y0), corner(x1, y1)
    return Ia * wa[:, None] + Ib * wb[:, None] + Ic * wc[:, None] + Id * wd[:, None]


def bilinear_sampler(img, coords):
    """``coords (N, Hg, Wg, 2)`` in pixels (x, y) -> ``(N, C, Hg, Wg)``.  Reference: utils.py:5-20."""
    img = np.asarray(img, dtype=F32)
    coords = np.asarray(coords, dtype=F32)
    N, C, H, W = img.shape
    _, Hg, Wg, _ = coords.shape
    x = pixel_round_trip(coords[..., 0], W).reshape(N, -1)
    y = pixel_round_trip(coords[..., 1], H).reshape(N, -1)
    return bilinear_sample_zero_pad(img, x, y).reshape(N, C, Hg, Wg)


def window_offsets(small_patch):
    """The nine integer (dx, dy) taps: 3x3 (dy outer, dx inner) or 1x9.  cost_volume.py:106-133 / :62-67."""
    if small_patch:
        return [(dx, dy) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
    return [(dx, 0) for dx in range(-4, 5)]


def replicate_pad(x, pady, padx):
    """Reference: ``manual_pad`` utils.py:29-31 (``F.pad(..., 'replicate')``)."""
    return np.pad(x, ((0, 0), (0, 0), (pady, pady), (padx, padx)), mode="edge")


def group_mean_dot(left, right_k):
    """``mean_c left[n,c,h,w] * right_k[n,c,h,w]`` per group of C/4 channels -> ``(N, 4, H, W)``."""
    N, C, H, W = left.shape
    prod = (left * right_k).reshape(N, NUM_GROUPS, C // NUM_GROUPS, H, W)
    return prod.astype(np.float64).mean(axis=2).astype(F32)


def corr_iter(fmap1, fmap2, flow, small_patch):
    """Iter-mode AGCL -> ``(N, 36, H, W)``, channel ``g*9 + k``.

    Reference: ``AGCL.corr_iter`` cost_volume.py:54-79 + ``get_correlation`` :28-52: the right map is
    first warped by the flow (zero-padded bilinear at ``grid + flow``), then *replicate*-padded, and the
    window taps index the warped map (h outer, w inner, :43-44).
    """
    L = np.asarray(fmap1, dtype=F32)
    R = np.asarray(fmap2, dtype=F32)
    flow = np.asarray(flow, dtype=F32)
    N, C, H, W = L.shape
    coords = (coords_grid(N, H, W) + flow).transpose(0, 2, 3, 1)
    Rw = bilinear_sampler(R, coords)
    taps = window_offsets(small_patch)
    pady = max(abs(dy) for _, dy in taps)
    padx = max(abs(dx) for dx, _ in taps)
    Rp = replicate_pad(Rw, pady, padx)
    out = np.empty((N, NUM_GROUPS, SEARCH_NUM, H, W), dtype=F32)
    for k, (dx, dy) in enumerate(taps):
        crop = Rp[:, :, pady + dy:pady + dy + H, padx + dx:padx + dx + W]
        out[:, :, k] = group_mean_dot(L, crop)
    return out.reshape(N, NUM_GROUPS * SEARCH_NUM, H, W)


def corr_att_offset(fmap1, fmap2, flow, extra_offset, small_patch, att=None):
    """Offset-mode AGCL -> ``(N, 36, H, W)``, channel ``g*9 + k``.

    Reference: ``AGCL.corr_att_offset`` cost_volume.py:81-154.  Sample k of pixel p is taken at
    ``(grid + flow)(p) + (d_k + off_k(p))`` -- in that association order (:135-138,:135) -- where
    ``off_k = (extra_offset[:, 2k], extra_offset[:, 2k+1])`` = (x, y) (:114).
    """
    L = np.asarray(fmap1, dtype=F32)
    R = np.asarray(fmap2, dtype=F32)
    flow = np.asarray(flow, dtype=F32)
    extra = np.asarray(extra_offset, dtype=F32)
    N, C, H, W = L.shape
    if att is not None:
        lt = L.transpose(0, 2, 3, 1).reshape(N, H * W, C)
        rt = R.transpose(0, 2, 3, 1).reshape(N, H * W, C)
        lt, rt = att(lt, rt)
        L = np.asarray(lt, dtype=F32).reshape(N, H, W, C).transpose(0, 3, 1, 2)
        R = np.asarray(rt, dtype=F32).reshape(N, H, W, C).transpose(0, 3, 1, 2)
    base = coords_grid(N, H, W) + flow                      # (N,2,H,W)
    extra = extra.reshape(N, SEARCH_NUM, 2, H, W)
    out = np.empty((N, NUM_GROUPS, SEARCH_NUM, H, W), dtype=F32)
    for k, (dx, dy) in enumerate(window_offsets(small_patch)):
        px = base[:, 0] + (F32(dx) + extra[:, k, 0])
        py = base[:, 1] + (F32(dy) + extra[:, k, 1])
        Rk = bilinear_sampler(R, np.stack([px, py], axis=-1))
        out[:, :, k] = group_mean_dot(L, Rk)
    return out.reshape(N, NUM_GROUPS * SEARCH_NUM, H, W)


class AGCL:
    """Oracle twin of cre_stereo/cost_volume.py:6-26."""

    def __init__(self, fmap1, fmap2, att=None):
        self.fmap1, self.fmap2, self.att = fmap1, fmap2, att

    def __call__(self, flow, extra_offset, small_patch=False, iter_mode=False):
        if iter_mode:
            return corr_iter(self.fmap1, self.fmap2, flow, small_patch)
        return corr_att_offset(self.fmap1, self.fmap2, flow, extra_offset, small_patch, self.att)

"""Oracle (test infrastructure, numpy): IGEV-Stereo group-wise volume, dual lookup, soft-argmin.

Restates ``nndepth/models/igev_stereo/cost_volume.py:9-98`` and the regression at
``nndepth/models/igev_stereo/model.py:92-95,144-146``.  The 3-D regulariser (``CostVolumeFilterNetwork``,
cost_volume.py:101-210: Conv3d/BatchNorm3d hourglass) is OUT OF SCOPE and enters only as a callable.
Pinned by ``tests/golden/igev_*.npz``.
"""
import numpy as np

from .corr1d import F32, avg_pool_pairs, level_positions, linear_sampler


def groupwise_volume(fmap1, fmap2, num_groups=8, accumulate="f64"):
    """``vol[b,g,h,w1,w2] = sum_{c in [G*g, G*g+G)} f1[b,c,h,w1] f2[b,c,h,w2] / G**0.5``.

    Reference: ``GeometryAwareCostVolume.build_cost_volume`` igev_stereo/cost_volume.py:81-98.
    ``torch.split(fmap, num_groups, dim=1)`` yields chunks of *size* ``num_groups`` and only the first
    ``num_groups`` chunks are consumed (:90), so only the first ``G*G`` channels matter (64 of 256)
    and each group's scale is ``1/sqrt(chunk size) = 1/sqrt(G)`` (:92,95).  All-pairs: D = W2.
    """
    f1 = np.asarray(fmap1, dtype=F32)
    f2 = np.asarray(fmap2, dtype=F32)
    G = num_groups
    if f1.shape[1] % G or f2.shape[1] % G:
        raise AssertionError("Number of channels of fmap1 and fmap2 must be the factor of num_groups")
    if f1.shape[1] < G * G:
        raise IndexError("the reference indexes chunk i of size G for i < G: needs C >= G*G")
    acc_t = np.float64 if accumulate == "f64" else F32
    vols = []
    for g in range(G):
        a = f1[:, G * g:G * g + G].astype(acc_t)
        b = f2[:, G * g:G * g + G].astype(acc_t)
        vols.append(np.einsum("bchi,bchj->bhij", a, b, optimize=True).astype(F32) / F32(G ** 0.5))
    return np.stack(vols, axis=1)


def volume_pyramids(feat_volume, geo_volume, num_levels=4):
    """Two ``num_levels+1``-level pyramids of ``(B*G*H*W1, w_l)`` rows.

    Reference: ``GeometryAwareCostVolume.__init__`` igev_stereo/cost_volume.py:39-52.
    ``feat_volume`` is ``(B,G,H,W1,W2)``; ``geo_volume`` is the regulariser output ``(B,G,W2,H,W1)``
    which is permuted back to ``(B,G,H,W1,W2)`` first (:45).
    """
    feat = np.asarray(feat_volume, dtype=F32)
    geo = np.asarray(geo_volume, dtype=F32).transpose(0, 1, 3, 4, 2)
    assert feat.shape == geo.shape
    fl = feat.reshape(-1, feat.shape[-1])
    gl = np.ascontiguousarray(geo).reshape(-1, geo.shape[-1])
    feat_pyr, geo_pyr = [fl], [gl]
    for _ in range(num_levels):
        fl = avg_pool_pairs(fl)
        gl = avg_pool_pairs(gl)
        feat_pyr.append(fl)
        geo_pyr.append(gl)
    return feat_pyr, geo_pyr


def gev_lookup(feat_pyr, geo_pyr, coords, num_levels=4, radius=4, num_groups=8):
    """Dual lookup -> ``(B, L*2*G*(2r+1), H, W)``; channel ``l*(2*G*T) + src*(G*T) + g*T + k``.

    Reference: ``GeometryAwareCostVolume.forward`` igev_stereo/cost_volume.py:54-79 (src 0 = feature
    correlation, src 1 = geometry volume; both sampled at the same positions).
    """
    coords = np.asarray(coords, dtype=F32)
    B, _, H, W = coords.shape
    G = num_groups
    rep = np.broadcast_to(coords.reshape(B, 1, H, W), (B, G, H, W)).reshape(-1)
    chunks = []
    for lvl in range(num_levels):
        x = level_positions(rep, lvl, radius)
        for pyr in (feat_pyr, geo_pyr):
            rows = pyr[lvl].reshape(B * G * H * W, -1)
            s = linear_sampler(rows, x).reshape(B, G, H, W, -1)
            chunks.append(s.transpose(0, 2, 3, 1, 4).reshape(B, H, W, -1))
    out = np.concatenate(chunks, axis=-1)
    return np.ascontiguousarray(out.transpose(0, 3, 1, 2), dtype=F32)


def softmax_disparity(z):
    """``softmax`` over axis 1 of ``(B, D, H, W)`` (igev_stereo/model.py:145), max-subtracted, fp32."""
    z = np.asarray(z, dtype=F32)
    e = np.exp(z - z.max(axis=1, keepdims=True))
    return e / e.sum(axis=1, keepdims=True, dtype=F32)


def regress_disparity(distribution, width):
    """``-sum_d d * p[b,d,h,w]`` -> ``(B,1,H,W)``.  Reference: igev_stereo/model.py:92-95."""
    p = np.asarray(distribution, dtype=F32)
    d = np.arange(width, dtype=F32).reshape(1, -1, 1, 1)
    return -np.sum(d * p, axis=1, keepdims=True, dtype=F32)


def soft_argmin(z, accumulate="f64"):
    """Fused softmax + expectation (model.py:145-146) -> ``(B,1,H,W)``; fp64 inside by default."""
    z = np.asarray(z, dtype=F32)
    acc_t = np.float64 if accumulate == "f64" else F32
    zz = z.astype(acc_t)
    e = np.exp(zz - zz.max(axis=1, keepdims=True))
    d = np.arange(z.shape[1], dtype=acc_t).reshape(1, -1, 1, 1)
    return (-(d * e).sum(axis=1, keepdims=True) / e.sum(axis=1, keepdims=True)).astype(F32)


def squeeze_cost(geo_level0, shape, weight, bias=None, accumulate="f64"):
    """``cv_squeezer`` = ``nn.Conv3d(G, 1, 3, 1, 1)`` on the level-0 geometry volume as the model feeds it.

    Reference: igev_stereo/model.py:65 (the layer) and :143-145 (``geo_aware_cv[0].reshape(B, G, H, W1, W2)
    .permute(0, 1, 4, 2, 3)`` -> conv -> ``squeeze(1)``).  ``geo_level0`` is the ``(B*G*H*W1, W2)`` level;
    ``shape = (B, G, H, W1, W2)``; ``weight`` ``(1, G, 3, 3, 3)`` over (d, h, w), zero padding.  -> ``(B, D, H, W1)``.
    """
    B, G, H, W1, W2 = (int(v) for v in shape)
    acc_t = np.float64 if accumulate == "f64" else F32
    x = np.asarray(geo_level0, dtype=F32).reshape(B, G, H, W1, W2).transpose(0, 1, 4, 2, 3).astype(acc_t)
    w = np.asarray(weight, dtype=F32).reshape(G, 3, 3, 3).astype(acc_t)
    xp = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1), (1, 1)))
    out = np.zeros((B, W2, H, W1), dtype=acc_t)
    for kd in range(3):
        for kh in range(3):
            for kw in range(3):
                win = xp[:, :, kd:kd + W2, kh:kh + H, kw:kw + W1]
                out += np.einsum("bgdhw,g->bdhw", win, w[:, kd, kh, kw])
    if bias is not None:
        out += acc_t(np.asarray(bias, dtype=F32).reshape(-1)[0])
    return out.astype(F32)


def squeeze_soft_argmin(geo_level0, shape, weight, bias=None, accumulate="f64"):
    """Initial disparity of IGEV-Stereo: ``regress_disparity(softmax(cv_squeezer(geo)))``, model.py:143-146."""
    return soft_argmin(squeeze_cost(geo_level0, shape, weight, bias, accumulate), accumulate)


class GeometryAwareCostVolume:
    """Oracle twin of igev_stereo/cost_volume.py:9-79.  ``regularizer_3d`` maps numpy -> numpy."""

    def __init__(self, fmap1, fmap2, features, regularizer_3d, num_levels=4, radius=4, num_groups=8,
                 accumulate="f64"):
        self.num_levels, self.radius, self.num_groups = num_levels, radius, num_groups
        feat = groupwise_volume(fmap1, fmap2, num_groups, accumulate)
        geo = regularizer_3d(np.ascontiguousarray(feat.transpose(0, 1, 4, 2, 3)), features)
        self.feat_corr_cv, self.geo_aware_cv = volume_pyramids(feat, geo, num_levels)

    def __call__(self, coords):
        return gev_lookup(self.feat_corr_cv, self.geo_aware_cv, coords, self.num_levels, self.radius,
                          self.num_groups)

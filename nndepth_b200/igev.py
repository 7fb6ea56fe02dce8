"""IGEV-Stereo geometry-aware cost volume and soft-argmin on the sm_100a kernels.

Mirrors ``nndepth/models/igev_stereo/cost_volume.py:9-98`` (``GeometryAwareCostVolume``) and the
regression at ``nndepth/models/igev_stereo/model.py:92-95,144-146`` of the reference.
``model.corr_fn = nndepth_b200.GeometryAwareCostVolume`` swaps the reference model onto these kernels
(``igev_stereo/model.py:64,133-141``).  The 3-D regulariser stays the caller's PyTorch module.

Layout: for the model's configuration (8 groups, ``W2 % 8 == 0``, <= 4 levels, radius 4) the two lookup
pyramids are stored pixel-major with the group innermost (``[b][h][w1][d][g]``, see csrc/igev.cu), which
makes the eight windows of a pixel one contiguous run; the reference-layout lists ``feat_corr_cv`` /
``geo_aware_cv`` are synthesised on demand.  Other configurations use the reference row layout.
"""
import math

import torch
import torch.nn as nn

from . import _lib
from .corr import PyramidStorage, _check_coords, _GroupedBuild, _GroupedLookup, _train_f32, _wants_grad


class InterleavedPyramid:
    """Levels ``[b][h][w1][d_l][g]`` of a grouped volume in ONE allocation (G = 8, D % 8 == 0)."""

    def __init__(self, pixels, depth, num_levels, device):
        self.pixels, self.depth = int(pixels), int(depth)
        self.widths = [self.depth >> l for l in range(num_levels)]
        sizes = [self.pixels * w * 8 for w in self.widths]
        self.buffer = torch.empty(sum(sizes), dtype=torch.float32, device=device)
        self.levels, start = [], 0
        for n in sizes:
            self.levels.append(self.buffer[start:start + n])
            start += n

    @property
    def _level_ptrs(self):
        return _lib.ptr_array(self.levels)

    def fill(self, src, layout, src_pitch, B, D, H, W1):
        _lib.ops().gev_interleave_pool(src, layout, src_pitch, B, D, H, W1, self.buffer, len(self.levels))
        return self

    def load_reference(self, levels, B, H, W1):
        """From reference-layout levels ``(B*8*H*W1, w_l)`` (parity tests)."""
        for dst, w, src in zip(self.levels, self.widths, levels):
            src = torch.as_tensor(src).to(device=dst.device, dtype=torch.float32).reshape(B, 8, H, W1, w)
            dst.copy_(src.permute(0, 2, 3, 4, 1).reshape(-1))
        return self

    def reference_level(self, l, B, H, W1):
        """Level ``l`` back in the reference layout ``(B*8*H*W1, 1, w_l)``."""
        w = self.widths[l]
        return self.levels[l].view(B, H, W1, w, 8).permute(0, 4, 1, 2, 3).reshape(B * 8 * H * W1, 1, w)


class _GeoPool(torch.autograd.Function):
    """Differentiable ``(B, G, D, H, W1)`` geometry volume -> row-layout pyramid (``nnd_geo_transpose_pool``, reference
    igev_stereo/cost_volume.py:44-52); backward = un-pool the level gradients and transpose level 0 back."""

    @staticmethod
    def forward(ctx, geo, num_levels):
        B, G, D, H, W1 = geo.shape
        pyr = PyramidStorage(B * G * H * W1, D, num_levels, geo.device)
        _lib.ops().geo_transpose_pool(geo, pyr.buffer, num_levels)
        ctx.geom = (B, G, D, H, W1, num_levels)
        return pyr.buffer

    @staticmethod
    def backward(ctx, d_buffer):
        B, G, D, H, W1, L = ctx.geom
        d = PyramidStorage(B * G * H * W1, D, L, d_buffer.device, buffer=d_buffer.contiguous().clone())
        _lib.ops().pyramid_unpool_(d.buffer, d.rows, D, L)
        return d.levels[0][:, :D].reshape(B, G, H, W1, D).permute(0, 1, 4, 2, 3).contiguous(), None


class GeometryAwareCostVolume(nn.Module):
    """Group-wise all-pairs volume + regularised geometry volume, pooled, with a fused dual lookup."""

    def __init__(self, fmap1, fmap2, features, regularizer_3d, num_levels=4, radius=4, num_groups=8):
        super().__init__()
        self.num_groups = num_groups
        self.num_levels = num_levels
        self.radius = radius
        self._graph = None
        train_f = _wants_grad(fmap1, fmap2)
        train = train_f or (torch.is_grad_enabled() and isinstance(regularizer_3d, nn.Module)
                            and any(p.requires_grad for p in regularizer_3d.parameters()))
        f1 = _train_f32(fmap1, "fmap1") if train_f else _lib.as_cuda_f32(fmap1, "fmap1")
        f2 = _train_f32(fmap2, "fmap2") if train_f else _lib.as_cuda_f32(fmap2, "fmap2")
        if f1.dim() != 4 or f1.shape[:3] != f2.shape[:3]:
            raise RuntimeError("fmap1 and fmap2 must be (B, C, H, W) with equal batch, channels and height")
        B, C, H, W1 = f1.shape
        W2 = f2.shape[3]
        G = num_groups
        self._shape = (B, H, W1, W2)
        if train:
            self._init_differentiable(f1, f2, features, regularizer_3d, train_f)
            return
        self._interleaved = (G == 8 and W2 % 8 == 0 and W1 % 4 == 0 and num_levels <= 4 and radius == 4)
        # feature volume in the reference row layout: level 0 only on the interleaved path (it feeds the
        # regulariser), all levels otherwise
        self._feat = PyramidStorage(B * G * H * W1, W2, 1 if self._interleaved else num_levels, f1.device)
        self._build_feature_volume(f1, f2, self._feat)
        # regulariser input: (B, G, W2, H, W1) view of the level-0 volume (cost_volume.py:37); level 0 is
        # dense whenever W2 % 4 == 0, otherwise the padding columns are sliced off.
        feat0 = self._feat.levels[0][:, :W2].reshape(B, G, H, W1, W2)
        geo = regularizer_3d(feat0.clone().permute(0, 1, 4, 2, 3), features)
        if geo.dim() != 5 or geo.shape[1] != G:
            raise AssertionError("N must be equal to num_groups")
        geo = _lib.as_cuda_f32(geo, "regularizer_3d output")
        Bg, _, D, Hg, Wg = geo.shape
        if (Bg, Hg, Wg, D) != (B, H, W1, W2):
            raise RuntimeError(
                f"regularizer_3d returned {tuple(geo.shape)}; expected (B, G, W2, H, W1) = {(B, G, W2, H, W1)}")
        if self._interleaved:
            self._feat_il = InterleavedPyramid(B * H * W1, W2, num_levels, f1.device).fill(
                self._feat.levels[0], 0, self._feat.pitches[0], B, W2, H, W1)
            self._geo_il = InterleavedPyramid(B * H * W1, D, num_levels, f1.device).fill(geo, 1, 0, B, D, H, W1)
            self._geo = None
            return
        self._geo = PyramidStorage(B * G * H * W1, D, num_levels, f1.device)
        _lib.ops().geo_transpose_pool(geo, self._geo.buffer, num_levels)

    def _init_differentiable(self, f1, f2, features, regularizer_3d, train_f):
        """Training: both pyramids in the reference row layout, every step an autograd node -- group-wise build
        (``_GroupedBuild``), the caller's 3-D regulariser (plain PyTorch), transpose + pool (``_GeoPool``); the dual
        lookup then differentiates through ``_GroupedLookup``.  Gradients reach the feature maps, the guide features
        and the regulariser's parameters exactly as in the reference (igev_stereo/cost_volume.py:36-52)."""
        B, H, W1, W2 = self._shape
        G, L = self.num_groups, self.num_levels
        C = f1.shape[1]
        assert C % G == 0 and f2.shape[1] % G == 0, \
            "Number of channels of fmap1 and fmap2 must be the factor of num_groups"
        if G * G > C:
            raise IndexError("tuple index out of range")
        self._interleaved = False
        rows = B * G * H * W1
        if train_f:
            buf_feat = _GroupedBuild.apply(f1, f2, G, G, math.sqrt(G), L)
        else:
            buf_feat = PyramidStorage(rows, W2, L, f1.device).buffer
            _lib.ops().groupcorr_build(f1, f2, buf_feat, G, G, float(math.sqrt(G)), L)
        self._feat = PyramidStorage(rows, W2, L, f1.device, buffer=buf_feat.detach())
        pitch0 = self._feat.pitches[0]
        feat0 = buf_feat[:rows * pitch0].view(rows, pitch0)[:, :W2].reshape(B, G, H, W1, W2)
        geo = regularizer_3d(feat0.clone().permute(0, 1, 4, 2, 3), features)
        if geo.dim() != 5 or geo.shape[1] != G:
            raise AssertionError("N must be equal to num_groups")
        Bg, _, D, Hg, Wg = geo.shape
        if (Bg, Hg, Wg, D) != (B, H, W1, W2):
            raise RuntimeError(
                f"regularizer_3d returned {tuple(geo.shape)}; expected (B, G, W2, H, W1) = {(B, G, W2, H, W1)}")
        geo = _train_f32(geo, "regularizer_3d output")
        if geo.requires_grad:
            buf_geo = _GeoPool.apply(geo, L)
        else:
            buf_geo = PyramidStorage(rows, D, L, f1.device).buffer
            _lib.ops().geo_transpose_pool(geo, buf_geo, L)
        self._geo = PyramidStorage(rows, D, L, f1.device, buffer=buf_geo.detach())
        self._graph = (buf_feat, buf_geo)

    @classmethod
    def from_pyramids(cls, feat_levels, geo_levels, batch, height, num_levels=4, radius=4, num_groups=8,
                      device="cuda"):
        """Wrap two existing pyramids (lists of ``(B*G*H*W1, w_l)`` arrays) -- used by the parity tests."""
        self = cls.__new__(cls)
        nn.Module.__init__(self)
        self._graph = None
        self.num_groups, self.num_levels, self.radius = num_groups, num_levels, radius
        first = torch.as_tensor(feat_levels[0])
        rows, W2 = first.reshape(first.shape[0], -1).shape
        self._shape = (batch, height, rows // (batch * height * num_groups), W2)
        dev = torch.device(device)
        W1 = self._shape[2]
        self._interleaved = (num_groups == 8 and W2 % 8 == 0 and W1 % 4 == 0 and num_levels <= 4 and radius == 4)
        if self._interleaved:
            self._feat = PyramidStorage(rows, W2, 1, dev).load(feat_levels[:1])
            self._geo = None
            self._feat_il = InterleavedPyramid(batch * height * W1, W2, num_levels, dev).load_reference(
                feat_levels[:num_levels], batch, height, W1)
            self._geo_il = InterleavedPyramid(batch * height * W1, W2, num_levels, dev).load_reference(
                geo_levels[:num_levels], batch, height, W1)
        else:
            self._feat = PyramidStorage(rows, W2, num_levels, dev).load(feat_levels[:num_levels])
            self._geo = PyramidStorage(rows, W2, num_levels, dev).load(geo_levels[:num_levels])
        return self

    def _build_feature_volume(self, f1, f2, pyr):
        B, C, H, W1 = f1.shape
        W2 = f2.shape[3]
        G = self.num_groups
        assert C % G == 0 and f2.shape[1] % G == 0, \
            "Number of channels of fmap1 and fmap2 must be the factor of num_groups"
        if G * G > C:
            raise IndexError("tuple index out of range")  # reference reads chunk i < G of size G (:90)
        _lib.ops().groupcorr_build(f1, f2, pyr.buffer, G, G, float(math.sqrt(G)), pyr.num_levels)

    def _reference_list(self, il):
        """The reference's ``num_levels + 1`` list from an interleaved pyramid (de-interleaved copies)."""
        B, H, W1, _ = self._shape
        views = [il.reference_level(l, B, H, W1) for l in range(len(il.levels))]
        views.append(torch.nn.functional.avg_pool1d(views[-1], 2))      # the level the reference never reads
        return views

    @property
    def feat_corr_cv(self):
        if self._graph is not None and torch.is_grad_enabled():
            return self._feat.graph_view(self._graph[0])
        return self._reference_list(self._feat_il) if self._interleaved else self._feat.reference_view()

    @property
    def geo_aware_cv(self):
        if self._graph is not None and torch.is_grad_enabled():
            return self._geo.graph_view(self._graph[1])      # the model reads geo_aware_cv[0] (igev_stereo/model.py:144)
        return self._reference_list(self._geo_il) if self._interleaved else self._geo.reference_view()

    def build_cost_volume(self, fmap1, fmap2):
        """``(B, G, H, W1, W2)`` group-wise volume (reference cost_volume.py:81-98)."""
        f1 = _lib.as_cuda_f32(fmap1, "fmap1")
        f2 = _lib.as_cuda_f32(fmap2, "fmap2")
        B, _, H, W1 = f1.shape
        W2 = f2.shape[3]
        pyr = PyramidStorage(B * self.num_groups * H * W1, W2, 1, f1.device)
        self._build_feature_volume(f1, f2, pyr)
        return pyr.levels[0][:, :W2].reshape(B, self.num_groups, H, W1, W2)

    def init_disparity(self, cv_squeezer, return_cost=False):
        """``regress_disparity(softmax(cv_squeezer(geo).squeeze(1)), W)`` -> ``(B, 1, H, W1)`` in one kernel.

        Fuses the ``nn.Conv3d(8, 1, 3, 1, 1)`` squeeze of the level-0 geometry volume with the soft-argmin
        (igev_stereo/model.py:65, 143-146): the volume is read once instead of permute-copied, convolved,
        soft-maxed and reduced.  ``cv_squeezer`` is the model's ``nn.Conv3d`` (weight ``(1, 8, 3, 3, 3)``).
        ``return_cost=True`` also returns the squeezed ``(B, D, H, W1)`` cost (parity tests).
        """
        B, H, W1, D = self._shape
        weight = _lib.as_cuda_f32(cv_squeezer.weight.detach(), "cv_squeezer.weight")
        bias = None if cv_squeezer.bias is None else _lib.as_cuda_f32(cv_squeezer.bias.detach(), "cv_squeezer.bias")
        if tuple(weight.shape) != (1, self.num_groups, 3, 3, 3):
            raise RuntimeError(f"cv_squeezer.weight must be (1, {self.num_groups}, 3, 3, 3), got {tuple(weight.shape)}")
        if tuple(cv_squeezer.padding) != (1, 1, 1) or tuple(cv_squeezer.stride) != (1, 1, 1) \
                or tuple(cv_squeezer.dilation) != (1, 1, 1):
            raise RuntimeError("cv_squeezer must be Conv3d(kernel 3, stride 1, padding 1, dilation 1)")
        if not self._interleaved or D > 512:
            # reference-layout volume: cuDNN Conv3d on the permuted view, then the fused soft-argmin kernel
            geo = self.geo_aware_cv[0].reshape(B, self.num_groups, H, W1, D).permute(0, 1, 4, 2, 3)
            cost = torch.nn.functional.conv3d(geo, weight, bias, 1, 1).squeeze(1).contiguous()
            disp = soft_argmin(cost)
            return (disp, cost) if return_cost else disp
        out, cost = _lib.ops().gev_squeeze_soft_argmin(self._geo_il.levels[0], weight, bias, B, self.num_groups, D, H, W1,
                                                       bool(return_cost))
        return (out, cost) if return_cost else out

    def forward(self, coords):
        B, H, W1, _ = self._shape
        if self._graph is not None and torch.is_grad_enabled():
            coords = _check_coords(coords.detach(), B, H, W1)     # the reference detaches them (igev_stereo/model.py:153)
            return _GroupedLookup.apply(self._graph[0], self._graph[1], coords, self._shape[3], self.num_groups,
                                        self.num_levels, self.radius, 0)
        coords = _check_coords(coords, B, H, W1)
        if self._interleaved:
            return _lib.ops().gev_lookup(self._feat_il.buffer, self._geo_il.buffer, coords, self._shape[3], self.num_levels,
                                         self.radius)
        return _lib.ops().group_lookup(self._feat.buffer, self._geo.buffer, self._shape[3], coords, self.num_groups,
                                       self.num_levels, self.radius, 0)


def soft_argmin(cost):
    """``-sum_d d * softmax_d(cost)``: ``(B, D, H, W)`` -> ``(B, 1, H, W)`` in one pass.

    Fuses ``F.softmax(dim=1)`` (igev_stereo/model.py:145) with ``regress_disparity`` (:92-95).
    """
    if torch.is_grad_enabled() and isinstance(cost, torch.Tensor) and cost.requires_grad:
        # training: the reference's own differentiable chain (softmax, then -sum d * p); the fused kernel is inference
        D = cost.shape[1]
        disp = torch.arange(0, D, dtype=cost.dtype, device=cost.device).reshape(1, -1, 1, 1)
        return -torch.sum(disp * torch.softmax(cost, dim=1), dim=1, keepdim=True)
    cost = _lib.as_cuda_f32(cost, "cost")
    if cost.dim() != 4:
        raise RuntimeError(f"cost must be (B, D, H, W), got {tuple(cost.shape)}")
    return _lib.ops().soft_argmin(cost)

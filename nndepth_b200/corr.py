"""RAFT-Stereo correlation blocks on the sm_100a kernels -- drop-in for the reference classes.

Mirrors ``nndepth/models/raft_stereo/cost_volume.py`` of the reference: ``CorrBlock1D`` (:7-61),
``GroupCorrBlock1D`` (:64-128) and ``linear_sampler`` (``raft_stereo/utils.py:4-27``).  Same
constructor / ``__call__`` signatures, attributes (``num_levels``, ``radius``, ``corr_pyramid``) and
output layout, so ``model.corr_fn = nndepth_b200.CorrBlock1D`` swaps the reference model onto these
kernels (``raft_stereo/model.py:58,124,132``).

Differences by design: one launch builds the volume *and* its pooled levels; one launch per GRU
iteration does the whole 4-level lookup; rows of the pyramid are padded to 16 bytes (``corr_pyramid``
hides the padding).  ``CorrBlock1D`` is differentiable with respect to the feature maps (the coordinates are
detached by the reference before every lookup, ``raft_stereo/model.py:131``): the lookup backward, the pooling
backward and the two volume-gradient contractions (``nnd_volume_grad``) are kernels of this package.  ``GroupCorrBlock1D`` and
the IGEV volume (``igev.py``) are differentiable the same way (grouped contraction, lookup backward per group).
"""
import math
import warnings

import torch

from . import _lib

_VOLUME_PRECISION = "tf32"


def set_volume_precision(precision):
    """``"tf32"`` (inference default: tcgen05 tensor cores, operands rounded to nearest TF32, 1e-3 bar) or
    ``"fp32"`` (CUDA-core FFMA, 1e-5 parity bar).  Two cases run on the fp32 kernel while the precision is left at
    its default: a differentiable build (feature maps that require grad, under grad mode -- the reference trains in
    fp32 unless the caller opts into autocast) and shapes the tensor-core path cannot take (feature widths not
    divisible by 4; warned once)."""
    global _VOLUME_PRECISION
    if precision not in ("fp32", "tf32"):
        raise ValueError(f"precision must be 'fp32' or 'tf32', got {precision!r}")
    _VOLUME_PRECISION = precision


def get_volume_precision():
    return _VOLUME_PRECISION


_warned = set()


def _warn_once(key, message):
    """One line per process and cause: a re-routed call must not be silent (there is no multi-backend dispatch)."""
    if key not in _warned:
        _warned.add(key)
        warnings.warn(message, RuntimeWarning, stacklevel=3)


def _prec_code(precision, W1=4, W2=4, train=False):
    if precision is None:
        precision = _VOLUME_PRECISION
        if precision == "tf32" and train:
            # the reference's training forward contracts in fp32 unless the caller opts into autocast: a drop-in
            # must not change the numerics of a training run behind the caller's back
            precision = "fp32"
        if precision == "tf32" and (W1 % 4 or W2 % 4):
            # 16-byte TMA rows need widths divisible by 4: same result (to 1e-5) on the slower fp32 FFMA kernel
            _warn_once(("width", W1, W2),
                       f"nndepth_b200: feature widths ({W1}, {W2}) are not multiples of 4; the correlation volume is "
                       "built by the fp32 CUDA-core kernel instead of the tcgen05 TF32 kernel (about 5x slower)")
            precision = "fp32"
    if precision not in ("fp32", "tf32"):
        raise ValueError(f"precision must be 'fp32' or 'tf32', got {precision!r}")
    return _lib.PREC_TF32 if precision == "tf32" else _lib.PREC_FP32


class PyramidStorage:
    """``num_levels`` pooled levels of ``rows`` volume rows in ONE device allocation.

    Level ``l`` is a ``(rows, pitch_l)`` fp32 matrix whose first ``width0 >> l`` columns are valid
    (``pitch_l`` = width rounded up to 4 floats so every row starts 16-byte aligned).
    """

    def __init__(self, rows, width0, num_levels, device, buffer=None):
        if not 1 <= num_levels <= _lib.NND_MAX_LEVELS:
            raise ValueError(f"num_levels must be in [1, {_lib.NND_MAX_LEVELS}], got {num_levels}")
        self.rows = int(rows)
        self.widths = [int(width0) >> l for l in range(num_levels)]
        if self.widths[-1] < 1:
            raise ValueError(f"a {num_levels}-level pyramid of width {width0} has an empty level")
        self.pitches = [_lib.row_pitch(w) for w in self.widths]
        self.buffer = (torch.empty(self.rows * sum(self.pitches), dtype=torch.float32, device=device)
                       if buffer is None else buffer)
        self.levels = []
        start = 0
        for p in self.pitches:
            self.levels.append(self.buffer[start:start + self.rows * p].view(self.rows, p))
            start += self.rows * p
        self._extra = None
        self.width0 = int(width0)

    # ctypes views of the level table, for callers that drive the C ABI directly (tests/test_abi.py, tools/)
    @property
    def _level_ptrs(self):
        return _lib.ptr_array(self.levels)

    @property
    def _width_arr(self):
        return _lib.int_array(self.widths)

    @property
    def _pitch_arr(self):
        return _lib.int_array(self.pitches)

    @property
    def num_levels(self):
        return len(self.levels)

    def load(self, levels):
        """Copy caller-provided levels (``(rows, w_l)`` or ``(rows, 1, w_l)`` tensors) into the padded storage."""
        for dst, w, src in zip(self.levels, self.widths, levels):
            src = torch.as_tensor(src).reshape(self.rows, -1)
            if src.shape[1] != w:
                raise RuntimeError(f"pyramid level has width {src.shape[1]}, expected {w}")
            dst[:, :w] = src.to(device=dst.device, dtype=torch.float32)
        return self

    def graph_view(self, graph_buffer):
        """The same list as views of an autograd-tracked copy of the buffer (training: gradients flow through them)."""
        views, start = [], 0
        for p, w in zip(self.pitches, self.widths):
            views.append(graph_buffer[start:start + self.rows * p].view(self.rows, p)[:, None, :w])
            start += self.rows * p
        views.append(torch.nn.functional.avg_pool1d(views[-1], 2))
        return views

    def reference_view(self):
        """The reference's list: ``num_levels + 1`` tensors ``(rows, 1, w_l)`` (cost_volume.py:29-34).

        The last level is never read by the reference's ``__call__``; it is pooled on first access.
        """
        views = [lv[:, None, :w] for lv, w in zip(self.levels, self.widths)]
        if self._extra is None:
            w_last = self.widths[-1]
            if w_last < 2:
                raise RuntimeError("avg_pool1d(kernel 2) of a 1-wide level: output size would be 0")
            self._extra = _lib.ops().avgpool_pairs(self.levels[-1], w_last)
        views.append(self._extra[:, None, :self.widths[-1] // 2])
        return views


class _BuildPyramid(torch.autograd.Function):
    """Differentiable pyramid build: forward = ``nnd_corr1d_build``; backward = un-pool the level gradients
    (``nnd_avgpool_pairs_backward``) and contract ``d_volume`` with the other feature map (``nnd_volume_grad``)."""

    @staticmethod
    def forward(ctx, f1, f2, num_levels, prec_code):
        B, C, H, W1 = f1.shape
        W2 = f2.shape[3]
        pyr = PyramidStorage(B * H * W1, W2, num_levels, f1.device)
        _lib.ops().corr1d_build(f1, f2, pyr.buffer, num_levels, prec_code)
        ctx.save_for_backward(f1, f2)
        ctx.geom = (B, C, H, W1, W2, num_levels)
        return pyr.buffer

    @staticmethod
    def backward(ctx, d_buffer):
        f1, f2 = ctx.saved_tensors
        B, C, H, W1, W2, L = ctx.geom
        d = PyramidStorage(B * H * W1, W2, L, f1.device, buffer=d_buffer.contiguous().clone())
        _lib.ops().pyramid_unpool_(d.buffer, d.rows, W2, L)     # avg_pool1d backward, coarsest level first
        d_f1 = _lib.ops().volume_grad(d.buffer, f2, W1, W2, 1, C, math.sqrt(C), 0) if ctx.needs_input_grad[0] else None
        d_f2 = _lib.ops().volume_grad(d.buffer, f1, W1, W2, 1, C, math.sqrt(C), 1) if ctx.needs_input_grad[1] else None
        return d_f1, d_f2, None, None


class _LookupPyramid(torch.autograd.Function):
    """Differentiable lookup: the gradient flows to the pyramid only (``nnd_corr1d_lookup_backward``)."""

    @staticmethod
    def forward(ctx, buffer, coords, block):
        ctx.block = block
        ctx.save_for_backward(coords)
        return block._lookup_raw(coords)

    @staticmethod
    def backward(ctx, grad_out):
        (coords,) = ctx.saved_tensors
        block = ctx.block
        B, H, W1, W2 = block._shape
        d_buffer = _lib.ops().corr1d_lookup_backward(grad_out.contiguous().float(), coords, W2, block.num_levels,
                                                     block.radius)
        return d_buffer, None, None


def _wants_grad(*tensors):
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)


def _train_f32(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError(f"{name} must be a CUDA tensor: nndepth_b200 has no CPU path")
    return t.float().contiguous()


class _GroupedBuild(torch.autograd.Function):
    """Differentiable group-wise pyramid build (rows ``[b][g][h][w1]``): forward = ``nnd_groupcorr_build``; backward =
    un-pool the level gradients (``nnd_avgpool_pairs_backward``) and contract ``d_volume`` group by group with the
    other feature map.  Only the first ``G * group_size`` channels take part (the reference's ``torch.split`` quirk)."""

    @staticmethod
    def forward(ctx, f1, f2, G, group_size, scale_div, num_levels):
        B, C, H, W1 = f1.shape
        W2 = f2.shape[3]
        pyr = PyramidStorage(B * G * H * W1, W2, num_levels, f1.device)
        _lib.ops().groupcorr_build(f1, f2, pyr.buffer, G, group_size, float(scale_div), num_levels)
        ctx.save_for_backward(f1, f2)
        ctx.geom = (B, C, H, W1, W2, G, group_size, float(scale_div), num_levels)
        return pyr.buffer

    @staticmethod
    def backward(ctx, d_buffer):
        f1, f2 = ctx.saved_tensors
        B, C, H, W1, W2, G, gs, scale_div, L = ctx.geom
        d = PyramidStorage(B * G * H * W1, W2, L, f1.device, buffer=d_buffer.contiguous().clone())
        _lib.ops().pyramid_unpool_(d.buffer, d.rows, W2, L)
        d_f1 = _lib.ops().volume_grad(d.buffer, f2, W1, W2, G, gs, scale_div, 0) if ctx.needs_input_grad[0] else None
        d_f2 = _lib.ops().volume_grad(d.buffer, f1, W1, W2, G, gs, scale_div, 1) if ctx.needs_input_grad[1] else None
        return d_f1, d_f2, None, None, None, None


class _GroupedLookup(torch.autograd.Function):
    """Differentiable grouped lookup over one or two row-layout pyramids (``nnd_group_lookup``).  The gradient flows to
    the pyramids only: per source it is the plain lookup backward (``nnd_corr1d_lookup_backward``) on ``B * G``
    "images" whose coordinates repeat per group, after undoing the output channel order of the mode
    (0: ``l*(S*G*T) + s*(G*T) + g*T + k``; 1: the ``GroupCorrBlock1D`` view quirk, reference cost_volume.py:108)."""

    @staticmethod
    def forward(ctx, buf_a, buf_b, coords, W2, G, num_levels, radius, mode):
        ctx.save_for_backward(coords)
        ctx.meta = (W2, G, num_levels, radius, mode, buf_b is not None)
        return _lib.ops().group_lookup(buf_a, buf_b, W2, coords, G, num_levels, radius, mode)

    @staticmethod
    def backward(ctx, grad_out):
        (coords,) = ctx.saved_tensors
        W2, G, L, r, mode, two = ctx.meta
        B, _, H, W1 = coords.shape
        T, S, hw = 2 * r + 1, 2 if two else 1, H * W1
        grad_out = grad_out.contiguous().float()
        coords_g = coords.repeat_interleave(G, dim=0).contiguous()          # (B*G, 1, H, W1): image b*G + g
        grads = []
        for s in range(S):
            if mode == 0:
                g = grad_out.view(B, L, S, G, T, H, W1)[:, :, s].permute(0, 2, 1, 3, 4, 5)      # (B, G, L, T, H, W1)
            else:
                # level block (GT, hw) is the transpose of the flat [g][rem][k] result, see lookup.cu mode 1
                g = grad_out.view(B, L, G * T, hw).transpose(2, 3).reshape(B, L, G, hw, T).permute(0, 2, 1, 4, 3)
            g = g.reshape(B * G, L * T, H, W1).contiguous()
            grads.append(_lib.ops().corr1d_lookup_backward(g, coords_g, W2, L, r))
        return grads[0], (grads[1] if two else None), None, None, None, None, None, None


def _check_coords(coords, B, H, W1):
    coords = _lib.as_cuda_f32(coords, "coords")
    if coords.dim() != 4 or coords.shape[1] != 1:
        raise RuntimeError(f"coords must be (B, 1, H, W), got {tuple(coords.shape)}")
    if tuple(coords.shape) != (B, 1, H, W1):
        raise RuntimeError(f"coords shape {tuple(coords.shape)} does not match the volume's (B,1,H,W1) = {(B, 1, H, W1)}")
    return coords


def _half_channels_last(fmap1, fmap2, precision):
    """True when both maps are dense fp16 channels-last CUDA tensors of a shape the tensor-core build takes and the
    volume precision in force is "tf32" (the fp16 products are its exact equivalent; "fp32" asks for the FFMA build)."""
    if (precision or _VOLUME_PRECISION) != "tf32":
        return False
    for t in (fmap1, fmap2):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float16 and t.dim() == 4
                and t.shape[1] > 1 and t.is_contiguous(memory_format=torch.channels_last)):
            return False
    return fmap1.shape[1] % 8 == 0 and fmap1.shape[3] % 4 == 0 and fmap2.shape[3] % 4 == 0


class CorrBlock1D:
    """All-pairs 1-D correlation pyramid + fused radius-r lookup (reference cost_volume.py:7-61)."""

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4, precision=None):
        self.num_levels = num_levels
        self.radius = radius
        self._graph_buffer = None
        train = torch.is_grad_enabled() and (getattr(fmap1, "requires_grad", False) or getattr(fmap2, "requires_grad", False))
        if train:
            # training: keep the autograd graph (the feature maps stay attached)
            for t, name in ((fmap1, "fmap1"), (fmap2, "fmap2")):
                if not (isinstance(t, torch.Tensor) and t.is_cuda):
                    raise RuntimeError(f"{name} must be a CUDA tensor: nndepth_b200 has no CPU path")
            f1, f2 = fmap1.float().contiguous(), fmap2.float().contiguous()
        elif _half_channels_last(fmap1, fmap2, precision):
            # fp16 channels-last maps (a cuDNN fp16 encoder's output) are contracted where they lie: no fp32 copy, no
            # NCHW copy, fp16 x fp16 products with fp32 accumulation (exactly the TF32 products of the same values)
            f1, f2 = fmap1.detach(), fmap2.detach()
        else:
            f1 = _lib.as_cuda_f32(fmap1, "fmap1")
            f2 = _lib.as_cuda_f32(fmap2, "fmap2")
        if f1.dim() != 4 or f2.dim() != 4:
            raise RuntimeError("fmap1 and fmap2 must be (B, C, H, W)")
        if f1.shape[:3] != f2.shape[:3] or f1.device != f2.device:
            raise RuntimeError(
                f"fmap1 {tuple(f1.shape)} and fmap2 {tuple(f2.shape)} must agree in batch, channels and height"
            )
        B, C, H, W1 = f1.shape
        W2 = f2.shape[3]
        self._shape = (B, H, W1, W2)
        if train:
            self._graph_buffer = _BuildPyramid.apply(f1, f2, num_levels, _prec_code(precision, W1, W2, train=True))
            self._pyr = PyramidStorage(B * H * W1, W2, num_levels, f1.device, buffer=self._graph_buffer.detach())
            return
        self._pyr = PyramidStorage(B * H * W1, W2, num_levels, f1.device)
        if f1.dtype == torch.float16:
            _lib.ops().corr1d_build_nhwc_f16(f1, f2, self._pyr.buffer, num_levels)
        else:
            _lib.ops().corr1d_build(f1, f2, self._pyr.buffer, num_levels, _prec_code(precision, W1, W2))

    @classmethod
    def from_pyramid(cls, levels, batch, height, num_levels=4, radius=4, device="cuda"):
        """Wrap an existing pyramid (list of ``(B*H*W1, w_l)`` arrays/tensors) -- used by the parity tests."""
        self = cls.__new__(cls)
        self.num_levels, self.radius = num_levels, radius
        self._graph_buffer = None
        first = torch.as_tensor(levels[0])
        rows, W2 = first.reshape(first.shape[0], -1).shape
        self._shape = (batch, height, rows // (batch * height), W2)
        self._pyr = PyramidStorage(rows, W2, num_levels, torch.device(device)).load(levels[:num_levels])
        return self

    @property
    def corr_pyramid(self):
        return self._pyr.reference_view()

    def __call__(self, coords, skewed=False):
        """``skewed=True`` (inference, radius 4) reads the skewed copy of the pyramid instead (``skewed_pyramid()``): the
        same values bit for bit, with a third of the DRAM traffic when the disparity field is smooth."""
        B, H, W1, _ = self._shape
        if self._graph_buffer is not None and torch.is_grad_enabled():
            coords = _check_coords(coords.detach(), B, H, W1)     # the reference detaches them too (model.py:131)
            return _LookupPyramid.apply(self._graph_buffer, coords, self)
        if skewed:
            if self.radius != 4:
                raise ValueError("the skewed lookup is built for radius 4")
            return _lib.ops().corr1d_lookup_skewed(self.skewed_pyramid(), self._shape[3], _check_coords(coords, B, H, W1),
                                                   self.num_levels, self.radius)
        return self._lookup_raw(_check_coords(coords, B, H, W1))

    def _lookup_raw(self, coords):
        return _lib.ops().corr1d_lookup(self._pyr.buffer, self._shape[3], coords, self.num_levels, self.radius)

    @staticmethod
    def prepare_conv1x1_weight(weight):
        """``(c_out, K[, 1, 1])`` conv weight -> the ``(K, c_out)`` k-major copy ``lookup_conv1x1`` consumes."""
        weight = _lib.as_cuda_f32(weight, "weight")
        return weight.reshape(weight.shape[0], -1).t().contiguous()

    def skewed_pyramid(self):
        """The skewed copy of the pyramid (``nnd_corr1d_skew``), built on first use: level ``l`` of an epipolar row is
        stored as ``S[j][w1]`` with ``j = ((w1 >> l) - w2) mod W2_l``, so the windows of neighbouring pixels with similar
        disparity share cache lines.  Read by ``__call__(coords, skewed=True)`` and ``lookup_conv1x1(..., skewed=True)``."""
        if getattr(self, "_skew", None) is None:
            B, H, W1, W2 = self._shape
            self._skew = _lib.ops().corr1d_skew(self._pyr.buffer, B, H, W1, W2, self.num_levels)
        return self._skew

    def lookup_conv1x1(self, coords, weight, bias=None, relu=True, weight_t=None, precision="fp32", channels_last=False,
                       half=False, skewed=False):
        """``relu(conv1x1(self(coords)))`` in one launch; the ``(B, L*(2r+1), H, W)`` lookup never reaches HBM.

        Fuses the lookup with the motion encoder's first layer (reference ``blocks/update_block.py:51,58``:
        ``cor = F.relu(self.convc1(corr))``).  ``weight`` is ``(c_out, L*(2r+1))`` or the conv's
        ``(c_out, L*(2r+1), 1, 1)``; returns ``(B, c_out, H, W)``.  Radius 4, 4 levels.  The kernel wants the
        weights k-major: pass ``weight_t=prepare_conv1x1_weight(weight)`` to transpose once per forward
        instead of once per call.  ``precision``: ``"fp32"`` (FFMA) or ``"tf32"`` (tensor cores, operands rounded
        to nearest TF32 -- what cuDNN does to this layer under ``allow_tf32``).
        """
        if precision not in ("fp32", "tf32"):
            raise ValueError(f"precision must be 'fp32' or 'tf32', got {precision!r}")
        if channels_last and precision != "tf32":
            raise ValueError("channels_last output is provided by the tensor-core ('tf32') path only")
        if half and not channels_last:
            raise ValueError("half=True (fp16 output) needs channels_last=True")
        B, H, W1, _ = self._shape
        coords = _check_coords(coords, B, H, W1)
        T = 2 * self.radius + 1
        if weight_t is None:
            weight_t = self.prepare_conv1x1_weight(weight)
        weight = _lib.require_cuda_f32(weight_t, "weight_t")
        if weight.dim() != 2 or weight.shape[0] != self.num_levels * T:
            raise RuntimeError(f"weight must have {self.num_levels * T} input channels, got {tuple(weight.shape)}")
        if bias is not None:
            bias = _lib.as_cuda_f32(bias, "bias")
        c_out = weight.shape[1]
        if not (channels_last and c_out <= 256):
            if half:
                raise ValueError("fp16 output needs c_out <= 256 (tensor-core path)")
            channels_last = False
        if skewed and channels_last and c_out == 256 and precision == "tf32" and self.num_levels == 4 and self.radius == 4:
            # same kernel, same bits, windows gathered from the skewed copy (smooth disparity fields: ~3x fewer DRAM bytes)
            return _lib.ops().corr1d_lookup_conv1x1_skewed(self.skewed_pyramid(), self._shape[3], coords, self.num_levels,
                                                           self.radius, weight, bias, bool(relu), 2 if half else 1)
        # channels-last results come back with NCHW shape and channels-last strides ((B, H, W, c_out) in memory)
        return _lib.ops().corr1d_lookup_conv1x1(self._pyr.buffer, self._shape[3], coords, self.num_levels, self.radius,
                                                weight, bias, bool(relu),
                                                _lib.PREC_TF32 if precision == "tf32" else _lib.PREC_FP32,
                                                (2 if half else 1) if channels_last else 0)

    def lookup_indices(self, coords):
        """Debug/parity helper: the int32 ``(idx0, idx1)`` of every tap, each ``(L, B*H*W1, 2r+1)``."""
        B, H, W1, _ = self._shape
        coords = _check_coords(coords, B, H, W1)
        return lookup_indices(self._pyr.widths, coords, self.num_levels, self.radius)

    @staticmethod
    def corr(fmap1, fmap2, precision=None):
        """``(B, H, W1, W2)`` volume only (reference cost_volume.py:55-61)."""
        f1 = _lib.as_cuda_f32(fmap1, "fmap1")
        f2 = _lib.as_cuda_f32(fmap2, "fmap2")
        B, C, H, W1 = f1.shape
        W2 = f2.shape[3]
        pyr = PyramidStorage(B * H * W1, W2, 1, f1.device)
        _lib.ops().corr1d_build(f1, f2, pyr.buffer, 1, _prec_code(precision, W1, W2))
        return pyr.levels[0][:, :W2].reshape(B, H, W1, W2)


def lookup_indices(widths, coords, num_levels=4, radius=4):
    """Integer window indices of ``linear_sampler`` (raft_stereo/utils.py:16-21) for every level/tap."""
    coords = _lib.require_cuda_f32(coords, "coords")
    return _lib.ops().corr1d_lookup_indices([int(w) for w in list(widths)[:num_levels]], coords, num_levels, radius)


def linear_sampler(corr, coords_lvl):
    """``(N, w2)`` rows sampled at ``(N, T)`` positions -> ``(N, T)`` (reference raft_stereo/utils.py:4-27).

    Runs the fused lookup kernel as a one-level, radius-0 lookup per tap column (the reference helper
    is only ever called through the correlation blocks; this standalone form exists for API parity).
    """
    corr = _lib.as_cuda_f32(corr, "corr")
    coords_lvl = _lib.as_cuda_f32(coords_lvl, "coords_lvl")
    if corr.dim() != 2 or coords_lvl.dim() != 2 or corr.shape[0] != coords_lvl.shape[0]:
        raise RuntimeError("linear_sampler expects corr (N, w2) and coords_lvl (N, T)")
    N, w2 = corr.shape
    T = coords_lvl.shape[1]
    pitch = _lib.row_pitch(w2)
    rows = corr
    if pitch != w2:
        rows = torch.zeros(N, pitch, dtype=torch.float32, device=corr.device)
        rows[:, :w2] = corr
    out = torch.empty(N, T, dtype=torch.float32, device=corr.device)
    flat = rows.reshape(-1)
    for t in range(T):
        col = coords_lvl[:, t].contiguous().view(1, 1, 1, N)
        out[:, t] = _lib.ops().corr1d_lookup(flat, w2, col, 1, 0).view(N)
    return out


class GroupCorrBlock1D:
    """Grouped correlation pyramid of ``Coarse2FineGroupRepViTRAFTStereo`` (reference cost_volume.py:64-128).

    Reproduces the reference's quirks: ``torch.split(fmap, num_groups)`` makes chunks *of size*
    ``num_groups`` and only the first ``num_groups`` chunks are read (:115-121); the scale is
    ``1/sqrt(C_total)`` (:125); the looked-up block ``[b][g][h][w][k]`` is reinterpreted as
    ``(B, H, W, G*(2r+1))`` (:108).  Any ``num_groups`` from 1 to 16 (every value with ``G * G <= C`` at C <= 256) and
    32 is built by its own instantiation of the kernel; other values raise ``NNDepthError`` (NND_ERR_UNSUPPORTED).
    """

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4, num_groups=4):
        self.num_levels = num_levels
        self.radius = radius
        self.num_groups = num_groups
        self._graph_buffer = None
        train = _wants_grad(fmap1, fmap2)
        f1 = _train_f32(fmap1, "fmap1") if train else _lib.as_cuda_f32(fmap1, "fmap1")
        f2 = _train_f32(fmap2, "fmap2") if train else _lib.as_cuda_f32(fmap2, "fmap2")
        if f1.dim() != 4 or f1.shape[:3] != f2.shape[:3]:
            raise RuntimeError("fmap1 and fmap2 must be (B, C, H, W) with equal batch, channels and height")
        B, C, H, W1 = f1.shape
        W2 = f2.shape[3]
        G = num_groups
        if G * G > C:
            raise IndexError("tuple index out of range")  # the reference indexes chunk i < G of size G
        self._shape = (B, H, W1, W2)
        if train:
            # training: the pyramid buffer stays in the autograd graph (the feature maps receive gradients)
            self._graph_buffer = _GroupedBuild.apply(f1, f2, G, G, math.sqrt(C), num_levels)
            self._pyr = PyramidStorage(B * G * H * W1, W2, num_levels, f1.device, buffer=self._graph_buffer.detach())
            return
        self._pyr = PyramidStorage(B * G * H * W1, W2, num_levels, f1.device)
        _lib.ops().groupcorr_build(f1, f2, self._pyr.buffer, G, G, float(math.sqrt(C)), num_levels)

    @classmethod
    def from_pyramid(cls, levels, batch, height, num_levels=4, radius=4, num_groups=4, device="cuda"):
        self = cls.__new__(cls)
        self.num_levels, self.radius, self.num_groups = num_levels, radius, num_groups
        self._graph_buffer = None
        first = torch.as_tensor(levels[0])
        rows, W2 = first.reshape(first.shape[0], -1).shape
        self._shape = (batch, height, rows // (batch * height * num_groups), W2)
        self._pyr = PyramidStorage(rows, W2, num_levels, torch.device(device)).load(levels[:num_levels])
        return self

    @property
    def corr_pyramid(self):
        if self._graph_buffer is not None and torch.is_grad_enabled():
            return self._pyr.graph_view(self._graph_buffer)
        return self._pyr.reference_view()

    def __call__(self, coords):
        B, H, W1, _ = self._shape
        if self._graph_buffer is not None and torch.is_grad_enabled():
            coords = _check_coords(coords.detach(), B, H, W1)     # the reference detaches them too (model.py:302)
            return _GroupedLookup.apply(self._graph_buffer, None, coords, self._shape[3], self.num_groups, self.num_levels,
                                        self.radius, 1)
        coords = _check_coords(coords, B, H, W1)
        return _lib.ops().group_lookup(self._pyr.buffer, None, self._shape[3], coords, self.num_groups, self.num_levels,
                                       self.radius, 1)

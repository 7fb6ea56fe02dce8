"""RAFT-Stereo model shell around the sm_100a correlation kernels.

The dense layers (feature encoder, motion encoder, separable ConvGRU, flow / mask heads) are OUT OF
SCOPE of the B200 hot path: they are cuDNN convolutions and stay
plain ``torch.nn`` modules.  This file only re-states their module tree so that

* parameter names / shapes equal the reference's ``BaseRAFTStereo`` (``nndepth/models/raft_stereo/
  model.py:17-163``, ``encoders/basic_encoder.py``, ``blocks/update_block.py``, ``blocks/gru.py``,
  ``blocks/residual_block.py``) -- reference checkpoints load with ``load_state_dict`` unchanged and a
  seeded construction draws the same random weights;
* ``forward(frame1, frame2) -> List[{"up_disp": (B,1,H,W)}]`` keeps the reference signature
  (model.py:111-139) while the correlation pyramid, the per-iteration lookup and the convex upsampling
  run on the kernels of this package, and the whole forward can be replayed as one CUDA graph.

``corr_fn`` is a class-valued attribute exactly as in the reference (model.py:58), so the same shell
runs the CPU oracle in tests / the CPU baseline (``model.corr_fn = <oracle class>``).
"""
import itertools

import torch
import torch.nn as nn
import torch.nn.functional as F

from .corr import CorrBlock1D, _warn_once
from .upsample import convex_upsample as fused_convex_upsample


class cudnn_tf32:
    """Context manager: let (or forbid) cuDNN / cuBLAS run fp32 convolutions on TF32 tensor cores."""

    def __init__(self, allow):
        self.allow = bool(allow)

    def __enter__(self):
        self.old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = self.allow
        torch.backends.cuda.matmul.allow_tf32 = self.allow

    def __exit__(self, *exc):
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = self.old
        return False


def rn_tf32(t):
    """Round an fp32 tensor to the nearest TF32 value (10-bit mantissa), ties away from zero."""
    return ((t.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def inference_weight(conv):
    """The convolution's weight as the tensor cores should see it: under ``allow_tf32`` rounded to the NEAREST
    TF32 value once (cuDNN's TF32 kernels truncate fp32 operands; with truncated weights the KITTI/32-iteration
    disparity sits 0.005 px from the reference, with rounded ones 0.003 px), otherwise untouched."""
    if not torch.backends.cudnn.allow_tf32:
        return conv.weight
    cache = getattr(conv, "_tf32_weight", None)
    if cache is None or cache[0] != conv.weight._version or cache[1].device != conv.weight.device:
        rounded = rn_tf32(conv.weight.detach())
        if conv.weight.is_contiguous(memory_format=torch.channels_last) and not conv.weight.is_contiguous():
            rounded = rounded.contiguous(memory_format=torch.channels_last)
        cache = (conv.weight._version, rounded)
        conv._tf32_weight = cache
    return cache[1]


def conv_relu(conv, x):
    """``relu(conv(x))``: on the GPU, inference, as cuDNN's fused convolution + bias + ReLU -- one kernel where
    ``F.relu(conv(x))`` costs three (convolution, bias add, clamp).  Same arithmetic, same TF32 policy."""
    if (x.is_cuda and not torch.is_grad_enabled() and conv.bias is not None and conv.padding_mode == "zeros"
            and isinstance(conv.padding, tuple) and hasattr(torch, "cudnn_convolution_relu")):
        return torch.cudnn_convolution_relu(x, inference_weight(conv), conv.bias, conv.stride, conv.padding, conv.dilation,
                                            conv.groups)
    return F.relu(conv(x))


def conv_plain(conv, x):
    """``conv(x)``; on the GPU at inference with the TF32-rounded weight of ``inference_weight``."""
    if x.is_cuda and not torch.is_grad_enabled() and conv.padding_mode == "zeros" and isinstance(conv.padding, tuple):
        return F.conv2d(x, inference_weight(conv), conv.bias, conv.stride, conv.padding, conv.dilation, conv.groups)
    return conv(x)


def _is_nhwc(t):
    return t.dim() == 4 and t.permute(0, 2, 3, 1).is_contiguous()


def _small_kernels_ok(t, half_ok=False):
    return t.is_cuda and (t.dtype == torch.float32 or (half_ok and t.dtype == torch.float16)) and not torch.is_grad_enabled()


def half_conv_params(conv):
    """``(weight, bias)`` of a convolution as channels-last fp16 (cached per weight version): fp16 carries TF32's
    10-bit mantissa, so this is the precision the TF32 path gives the weights, at twice the tensor-core rate."""
    key = (conv.weight._version, None if conv.bias is None else conv.bias._version, conv.weight.device)
    cache = getattr(conv, "_f16_params", None)
    if cache is None or cache[0] != key:
        w = conv.weight.detach().clamp(-65504.0, 65504.0).half().contiguous(memory_format=torch.channels_last)
        b = None if conv.bias is None else conv.bias.detach().clamp(-65504.0, 65504.0).half().contiguous()
        cache = (key, (w, b))
        conv._f16_params = cache
    return cache[1]


def _split16(w, b):
    """``(w_hi, w_lo, b_hi, b_lo)``: fp16 value + fp16 remainder of an fp32 weight / bias (weights good to 2**-22
    relative, down to fp16's subnormal spacing) -- the two-term form the tensor cores can consume exactly."""
    w = w.detach().float().clamp(-65504.0, 65504.0)
    w_hi = w.half()
    w_lo = (w - w_hi.float()).half()
    cl = torch.channels_last
    if b is None:
        return w_hi.contiguous(memory_format=cl), w_lo.contiguous(memory_format=cl), None, None
    b = b.detach().float().clamp(-65504.0, 65504.0)
    b_hi = b.half()
    return (w_hi.contiguous(memory_format=cl), w_lo.contiguous(memory_format=cl), b_hi.contiguous(),
            (b - b_hi.float()).half().contiguous())


def half_split_conv_params(conv):
    """``_split16`` of a convolution's parameters, cached per parameter version / device."""
    key = (conv.weight._version, None if conv.bias is None else conv.bias._version, conv.weight.device)
    cache = getattr(conv, "_f16_split", None)
    if cache is None or cache[0] != key:
        cache = (key, _split16(conv.weight, conv.bias))
        conv._f16_split = cache
    return cache[1]


def conv_add_relu_split16(x16, params, stride, padding, dilation=(1, 1), groups=1):
    """``relu(conv(x, w_hi + w_lo) + b)`` on fp16 tensor cores with fp32 accumulation: the remainder product is a plain
    fp16 convolution (its output is ~2**-11 of the result, so its own fp16 rounding is ~2**-22), the main product is
    cuDNN's fused convolution + add + bias + ReLU reading it.  A rounded WEIGHT perturbs every GRU iteration the same
    way (the error accumulates coherently over 32 iterations); rounded ACTIVATIONS are fresh noise each iteration."""
    w_hi, w_lo, b_hi, _ = params
    # no bias on the remainder product: PyTorch would add it with a separate elementwise kernel, and the fp16 rounding
    # of the BIAS (one value per channel, ~2**-11 of a number ~1/sqrt(fan_in)) is two orders of magnitude below the
    # accumulated weight rounding this split removes
    z = F.conv2d(x16, w_lo, None, stride, padding, dilation, groups)
    return torch.cudnn_convolution_add_relu(x16, w_hi, z, 1.0, b_hi, stride, padding, dilation, groups)


def _dither16(t, k, K):
    """fp16 rounding of ``t + ((k + 1/2) / K - 1/2) * ulp16(t)``: K roundings of the same fp32 tensor whose mean is within
    ulp / (2K) of it.  Used for weights inside the refinement loop: iteration ``i`` takes variant ``i mod K``, so the
    rounding error of a weight is no longer the same perturbation in all 32 iterations (which accumulates coherently,
    see ``conv_add_relu_split16``) but changes sign from one iteration to the next -- at the cost of ONE product."""
    t = t.detach().float().clamp(-65504.0, 65504.0)
    ulp = torch.exp2(torch.floor(torch.log2(t.abs().clamp_min(2.0 ** -14))) - 10.0)
    return (t + ((k + 0.5) / K - 0.5) * ulp).clamp(-65504.0, 65504.0).half()


def half_dither_conv_params(conv, k, K, weight=None, bias=None, tag="_f16_dither"):
    """Variant ``k`` of ``K`` of a convolution's (channels-last fp16 weight, fp16 bias), cached per parameter version."""
    key = (conv.weight._version, None if conv.bias is None else conv.bias._version, conv.weight.device, K)
    cache = getattr(conv, tag, None)
    if cache is None or cache[0] != key:
        w = conv.weight if weight is None else weight
        b = conv.bias if bias is None else bias
        cache = (key, [(_dither16(w, j, K).contiguous(memory_format=torch.channels_last),
                        None if b is None else _dither16(b, j, K).contiguous()) for j in range(K)])
        setattr(conv, tag, cache)
    return cache[1][k % K]


def conv_relu_f16(conv, x16):
    """``relu(conv(x))`` as cuDNN's fused fp16 convolution (fp32 accumulation) on a channels-last fp16 ``x``.  Weight
    form (set by ``RAFTStereo._set_exact16``): ``conv.exact16`` -- two-term ``w_hi + w_lo`` (two products);
    ``conv.dither16 = (k, K)`` -- variant ``k`` of ``K`` dithered roundings (one product); else plain fp16."""
    if getattr(conv, "exact16", False):
        return conv_add_relu_split16(x16, half_split_conv_params(conv), conv.stride, conv.padding, conv.dilation, conv.groups)
    dither = getattr(conv, "dither16", None)
    w, b = half_dither_conv_params(conv, *dither) if dither else half_conv_params(conv)
    return torch.cudnn_convolution_relu(x16, w, b, conv.stride, conv.padding, conv.dilation, conv.groups)


def flow_conv7x7_relu(conv, flow, half=False):
    """``relu(conv(flow))`` for the one-channel 7x7 ``convf1`` (reference blocks/update_block.py:53,60) as one fp32
    kernel writing channels-last (``half``: rounded to fp16 on the way out); anything else goes to cuDNN."""
    if not (_small_kernels_ok(flow) and flow.shape[1] == 1 and conv.kernel_size == (7, 7) and conv.padding == (3, 3)
            and conv.stride == (1, 1) and conv.dilation == (1, 1) and conv.groups == 1 and conv.bias is not None
            and conv.out_channels % 4 == 0 and 256 % (conv.out_channels // 4) == 0):
        if flow.is_cuda and not torch.is_grad_enabled():
            _warn_once("flow_conv7x7", "nndepth_b200: flow_conv7x7_relu shape outside the fused kernel's (1 -> C, 7x7, "
                       f"pad 3, C % 4 == 0): conv {tuple(conv.weight.shape)} runs on cuDNN instead")
        out = conv_relu(conv, flow)
        return out.half().contiguous(memory_format=torch.channels_last) if half else out
    from . import _lib
    N, _, H, W = flow.shape
    flow = flow.contiguous()
    cache = getattr(conv, "_tap_major_weight", None)
    if cache is None or cache[0] != conv.weight._version or cache[1].device != flow.device:
        cache = (conv.weight._version, conv.weight.detach().reshape(conv.out_channels, 49).t().contiguous())
        conv._tap_major_weight = cache
    weight = cache[1]                                # (49, Cout)
    out = torch.empty(N, H, W, conv.out_channels, dtype=torch.float16 if half else torch.float32, device=flow.device)
    with torch.cuda.device(flow.device):
        _lib.check(_lib.load().nnd_flow_conv7x7_relu(_lib.ptr(flow), _lib.ptr(weight), _lib.ptr(conv.bias.detach()), N, H, W,
                                                     conv.out_channels, _lib.ptr(out), int(half), _lib.stream_ptr(flow)),
                   "nnd_flow_conv7x7_relu")
    return out.permute(0, 3, 1, 2)


def nhwc_cat_f16(a, b):
    """``torch.cat([a, b], 1)`` of two channels-last maps (fp32 or fp16 each) as one channels-last fp16 tensor."""
    from . import _lib
    N, Ca, H, W = a.shape
    Cb = b.shape[1]
    if not (_is_nhwc(a) and _is_nhwc(b) and Ca % 4 == 0 and Cb % 4 == 0 and a.is_cuda):
        return torch.cat([a.half(), b.half()], 1).contiguous(memory_format=torch.channels_last)
    out = torch.empty(N, H, W, Ca + Cb, dtype=torch.float16, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(_lib.load().nnd_nhwc_cat_f16(_lib.ptr(a), int(a.dtype == torch.float16), Ca, _lib.ptr(b),
                                                int(b.dtype == torch.float16), Cb, N * H * W, _lib.ptr(out), _lib.stream_ptr(a)),
                   "nnd_nhwc_cat_f16")
    return out.permute(0, 3, 1, 2)


def flow_head_tail(conv, x, coords=None, org=None):
    """``conv(x)`` for the flow head's 3x3 ``C -> 1`` convolution (reference blocks/update_block.py:23,36) on a
    channels-last ``x``, as one fp32 kernel.  With ``coords`` / ``org`` it also performs the loop's update
    (raft_stereo/model.py:132-134) and returns ``(coords + delta, coords + delta - org)``; otherwise ``delta``.
    Returns ``None`` when the shape is not the kernel's (the caller then uses cuDNN)."""
    if not (_small_kernels_ok(x, half_ok=True) and conv.out_channels == 1 and conv.in_channels in (128, 256, 512) and _is_nhwc(x)
            and conv.kernel_size == (3, 3) and conv.padding == (1, 1) and conv.stride == (1, 1)
            and conv.dilation == (1, 1) and conv.groups == 1):
        if x.is_cuda and not torch.is_grad_enabled() and conv.out_channels == 1:
            _warn_once("flow_head_tail", "nndepth_b200: flow_head_tail shape outside the fused kernel's (channels-last "
                       f"128/256/512 -> 1, 3x3): conv {tuple(conv.weight.shape)} runs on cuDNN instead")
        return None
    from . import _lib
    N, C, H, W = x.shape
    cache = getattr(conv, "_nchw_weight", None)
    if cache is None or cache[0] != conv.weight._version or cache[1].device != x.device:
        cache = (conv.weight._version, conv.weight.detach().contiguous(memory_format=torch.contiguous_format).clone())
        conv._nchw_weight = cache
    weight = cache[1]
    bias = None if conv.bias is None else conv.bias.detach()
    lib = _lib.load()
    with torch.cuda.device(x.device):
        if coords is None:
            delta = torch.empty(N, 1, H, W, dtype=torch.float32, device=x.device)
            _lib.check(lib.nnd_flow_head_tail(_lib.ptr(x), int(x.dtype == torch.float16), _lib.ptr(weight),
                                              _lib.ptr(bias) if bias is not None else None, N, C,
                                              H, W, _lib.ptr(delta), None, None, None, None, _lib.stream_ptr(x)),
                       "nnd_flow_head_tail")
            return delta
        coords, org = coords.contiguous(), org.contiguous()
        new_coords, new_flow = torch.empty_like(coords), torch.empty_like(coords)
        _lib.check(lib.nnd_flow_head_tail(_lib.ptr(x), int(x.dtype == torch.float16), _lib.ptr(weight),
                                          _lib.ptr(bias) if bias is not None else None, N, C, H,
                                          W, None, _lib.ptr(coords), _lib.ptr(org), _lib.ptr(new_coords), _lib.ptr(new_flow),
                                          _lib.stream_ptr(x)), "nnd_flow_head_tail")
    return new_coords, new_flow


def fold_bn(conv, bn, tf32=False):
    """Weights and bias of ``bn(conv(x))`` as ONE convolution (eval-mode BatchNorm is affine per channel).

    ``tf32``: the weights are additionally rounded to nearest TF32 -- tensor-core kernels that take fp32
    operands truncate them, and a truncated weight is biased where a rounded one is not."""
    scale = bn.weight.detach() / torch.sqrt(bn.running_var.detach() + bn.eps)
    w = (conv.weight.detach() * scale.view(-1, 1, 1, 1)).contiguous()
    b = ((conv.bias.detach() if conv.bias is not None else 0.0) - bn.running_mean.detach()) * scale + bn.bias.detach()
    return (rn_tf32(w) if tf32 else w), b.contiguous()


def _fold_f16(conv, bn):
    """``fold_bn`` with the folded weight / bias as channels-last fp16 (fp16 carries TF32's mantissa)."""
    w, b = fold_bn(conv, bn, False)
    return (w.clamp(-65504.0, 65504.0).half().contiguous(memory_format=torch.channels_last),
            b.clamp(-65504.0, 65504.0).half().contiguous())


def _fold_key(mode, *pairs):
    """Cache key of a folded conv + BatchNorm chain: in-place weight updates (load_state_dict, .to()) bump ``_version``."""
    key = [mode]
    for conv, bn in pairs:
        key += [conv.weight._version, conv.weight.device, None if conv.bias is None else conv.bias._version,
                bn.weight._version, bn.bias._version, bn.running_mean._version, bn.running_var._version]
    return tuple(key)


def _fused_ok(x, *norms):
    """Inference on the GPU with eval-mode BatchNorm: the conv + BN (+ add) + ReLU chains run as single cuDNN calls."""
    return (x.is_cuda and not torch.is_grad_enabled() and hasattr(torch, "cudnn_convolution_relu")
            and all(isinstance(n, nn.BatchNorm2d) and not n.training and n.track_running_stats for n in norms))


def _norm(kind, planes):
    if kind == "batch":
        return nn.BatchNorm2d(planes)
    if kind == "instance":
        return nn.InstanceNorm2d(planes, affine=False)
    if kind == "group":
        return nn.GroupNorm(num_groups=planes // 8, num_channels=planes)
    if kind == "none":
        return nn.Sequential()
    raise AssertionError(f"norm_fn must be in group, batch, instance, or none, found {kind}")


class ResidualBlock(nn.Module):
    """Two 3x3 convs + always-on 1x1 projection shortcut (reference blocks/residual_block.py:6-63)."""

    def __init__(self, in_planes, planes, norm_fn="group", stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(in_planes, planes, 3, stride=stride, padding=1)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1)
        self.relu = nn.ReLU(inplace=True)
        self.norm1, self.norm2, self.norm3 = (_norm(norm_fn, planes) for _ in range(3))
        self.downsample = nn.Sequential(nn.Conv2d(in_planes, planes, 1, stride=stride), self.norm3)

    def _folded(self, half=False):
        """(w, b) of conv1+norm1, conv2+norm2, shortcut+norm3, folded once (inference)."""
        tf32 = "f16" if half else bool(torch.backends.cudnn.allow_tf32)
        key = _fold_key(tf32, (self.conv1, self.norm1), (self.conv2, self.norm2), (self.downsample[0], self.norm3))
        cache = getattr(self, "_fold_cache", None)
        if cache is None or cache[0] != key:
            if half:
                cache = (key, (_fold_f16(self.conv1, self.norm1), _fold_f16(self.conv2, self.norm2),
                               _fold_f16(self.downsample[0], self.norm3)))
            else:
                cache = (key, (fold_bn(self.conv1, self.norm1, tf32), fold_bn(self.conv2, self.norm2, tf32),
                               fold_bn(self.downsample[0], self.norm3, tf32)))
            self._fold_cache = cache
        return cache[1]

    def _folded_split16(self):
        key = _fold_key("f16s", (self.conv1, self.norm1), (self.conv2, self.norm2), (self.downsample[0], self.norm3))
        cache = getattr(self, "_fold_cache16s", None)
        if cache is None or cache[0] != key:
            cache = (key, tuple(_split16(*fold_bn(c, n, False))
                                for c, n in ((self.conv1, self.norm1), (self.conv2, self.norm2), (self.downsample[0], self.norm3))))
            self._fold_cache16s = cache
        return cache[1]

    def forward(self, x):
        if _fused_ok(x, self.norm1, self.norm2, self.norm3):
            # BatchNorm folded into the convolutions; conv + bias + ReLU and conv + bias + add + ReLU are single
            # cuDNN calls (the reference chain is conv, bias add, batch norm, clamp: four passes per layer)
            c1, c3 = self.conv1, self.downsample[0]
            if x.dtype == torch.float16 and getattr(self, "exact16", False):
                p1, p2, p3 = self._folded_split16()
                y = conv_add_relu_split16(x, p1, c1.stride, c1.padding, c1.dilation)
                y = conv_add_relu_split16(y, p2, self.conv2.stride, self.conv2.padding, self.conv2.dilation)
                y = y + F.conv2d(x, p3[1], None, c3.stride, c3.padding, c3.dilation)
                return torch.cudnn_convolution_add_relu(x, p3[0], y, 1.0, p3[2], c3.stride, c3.padding, c3.dilation, 1)
            (w1, b1), (w2, b2), (w3, b3) = self._folded(half=x.dtype == torch.float16)
            y = torch.cudnn_convolution_relu(x, w1, b1, c1.stride, c1.padding, c1.dilation, 1)
            y = torch.cudnn_convolution_relu(y, w2, b2, self.conv2.stride, self.conv2.padding, self.conv2.dilation, 1)
            return torch.cudnn_convolution_add_relu(x, w3, y, 1.0, b3, c3.stride, c3.padding, c3.dilation, 1)
        y = self.relu(self.norm1(self.conv1(x)))
        y = self.relu(self.norm2(self.conv2(y)))
        return self.relu(self.downsample(x) + y)


class BasicEncoder(nn.Module):
    """Stride-8 residual feature encoder (reference encoders/basic_encoder.py:8-93)."""

    def __init__(self, output_dim=128, norm_fn="batch", dropout=0.0):
        super().__init__()
        self.norm_fn = norm_fn
        self.norm1 = nn.GroupNorm(8, 64) if norm_fn == "group" else _norm(norm_fn, 64)
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3)
        self.relu1 = nn.ReLU(inplace=True)
        widths, strides, planes = (64, 96, 128), (1, 2, 2), 64
        for i, (w, s) in enumerate(zip(widths, strides), start=1):
            setattr(self, f"layer{i}", nn.Sequential(ResidualBlock(planes, w, norm_fn, s),
                                                     ResidualBlock(w, w, norm_fn, 1)))
            planes = w
        self.conv2 = nn.Conv2d(128, output_dim, 1)
        self.dropout = nn.Dropout2d(dropout) if dropout > 0 else None
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, (nn.BatchNorm2d, nn.InstanceNorm2d, nn.GroupNorm)):
                if m.weight is not None:
                    nn.init.constant_(m.weight, 1)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    def forward(self, x):
        pair = isinstance(x, (tuple, list))
        if pair:
            x = torch.cat(x, dim=0)
        half = bool(getattr(self, "half_convs", False)) and _fused_ok(x, self.norm1)
        if half:
            x = x.half()
        if _fused_ok(x, self.norm1):
            tf32 = "f16" if half else bool(torch.backends.cudnn.allow_tf32)
            key = _fold_key(tf32, (self.conv1, self.norm1))
            if getattr(self, "_fold_cache", None) is None or self._fold_cache[0] != key:
                w, b = _fold_f16(self.conv1, self.norm1) if half else fold_bn(self.conv1, self.norm1, tf32)
                # a fourth, all-zero input channel: with 3 channels cuDNN has no Blackwell kernel for this layer and
                # falls back to an sm80 one (763 us at KITTI x 16 images); 4 channels are TMA-addressable
                w = F.pad(w, (0, 0, 0, 0, 0, (-w.shape[1]) % 4))
                if self.conv1.weight.is_contiguous(memory_format=torch.channels_last) and not self.conv1.weight.is_contiguous():
                    w = w.contiguous(memory_format=torch.channels_last)
                self._fold_cache = (key, (w, b))
            w, b = self._fold_cache[1]
            if w.shape[1] != x.shape[1]:
                fmt = torch.channels_last if (x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous()) \
                    else torch.contiguous_format
                x = F.pad(x, (0, 0, 0, 0, 0, w.shape[1] - x.shape[1])).contiguous(memory_format=fmt)
            if half and getattr(self, "exact16", False):
                key = _fold_key("f16s", (self.conv1, self.norm1))
                if getattr(self, "_fold_cache16s", None) is None or self._fold_cache16s[0] != key:
                    wf, bf = fold_bn(self.conv1, self.norm1, False)
                    self._fold_cache16s = (key, _split16(F.pad(wf, (0, 0, 0, 0, 0, (-wf.shape[1]) % 4)), bf))
                x = conv_add_relu_split16(x, self._fold_cache16s[1], self.conv1.stride, self.conv1.padding, self.conv1.dilation)
            else:
                x = torch.cudnn_convolution_relu(x, w, b, self.conv1.stride, self.conv1.padding, self.conv1.dilation, 1)
        else:
            x = self.relu1(self.norm1(self.conv1(x)))
        x = self.layer3(self.layer2(self.layer1(x)))
        if x.dtype == torch.float16 and getattr(self, "exact16", False):
            w_hi, w_lo, b_hi, b_lo = half_split_conv_params(self.conv2)
            x = F.conv2d(x, w_hi, b_hi, self.conv2.stride, self.conv2.padding) + F.conv2d(x, w_lo, None, self.conv2.stride, self.conv2.padding)
        elif x.dtype == torch.float16:
            w16, b16 = half_conv_params(self.conv2)
            x = F.conv2d(x, w16, b16, self.conv2.stride, self.conv2.padding)
        else:
            x = conv_plain(self.conv2, x)
        if self.dropout is not None:
            x = self.dropout(x)
        return torch.split(x, x.shape[0] // 2, dim=0) if pair else x


class SepConvGRU(nn.Module):
    """Horizontal (1x5) then vertical (5x1) conv GRU (reference blocks/gru.py:5-37).

    Inference fusion: the z and r gates read the same input, so their two convolutions run as one
    conv with concatenated output channels (weights concatenated on the fly -- no new parameters).
    """

    def __init__(self, hidden_dim=128, input_dim=192 + 128):
        super().__init__()
        cin = hidden_dim + input_dim
        for tag, k, p in (("1", (1, 5), (0, 2)), ("2", (5, 1), (2, 0))):
            for gate in "zrq":
                setattr(self, f"conv{gate}{tag}", nn.Conv2d(cin, hidden_dim, k, padding=p))
        self._fused = {}
        # set by RAFTStereo.dense_precision: "fp32" = cuDNN fp32 convolutions for the recurrence ("mixed"),
        # "wsplit" = TF32 activations x split fp32 weights on tensor cores ("mixed2x"), None = the caller's flags
        self.recurrence = None
        self._split_w = {}

    def fuse_gates(self):
        """Enable the concatenated z|r gate convolution for inference (call any time; weights may still change)."""
        self._fuse_zr = True
        self._fused = {}

    def _fused_zr(self, tag):
        """Concatenated z|r weights of one half-step, keyed on the parameters' versions and device: in-place updates
        (``load_state_dict``, ``.to()``, an optimiser step) bump ``_version``, so a stale snapshot is never served."""
        cz, cr = getattr(self, f"convz{tag}"), getattr(self, f"convr{tag}")
        stamp = tuple(p._version for c in (cz, cr) for p in (c.weight, c.bias)) + (cz.weight.device, cz.weight.dtype)
        hit = self._fused.get(tag)
        if hit is None or hit[3] != stamp:
            hit = (torch.cat([cz.weight, cr.weight], 0).detach().contiguous(),
                   torch.cat([cz.bias, cr.bias], 0).detach().contiguous(), cz.padding, stamp)
            self._fused[tag] = hit
        return hit[:3]

    def _half_step(self, h, x, tag):
        hx = torch.cat([h, x], dim=1)
        if getattr(self, "_fuse_zr", False) and not torch.is_grad_enabled():
            # inference only: under grad mode the per-gate convolutions below keep convz / convr in the autograd graph
            w, b, pad = self._fused_zr(tag)
            z, r = torch.sigmoid(F.conv2d(hx, w, b, padding=pad)).chunk(2, dim=1)
        else:
            z = torch.sigmoid(getattr(self, f"convz{tag}")(hx))
            r = torch.sigmoid(getattr(self, f"convr{tag}")(hx))
        q = torch.tanh(getattr(self, f"convq{tag}")(torch.cat([r * h, x], dim=1)))
        return (1 - z) * h + z * q

    # ---- weight-split TF32: conv(x, w) ~ conv(RN(x), w_hi) + conv(RN(x), w_lo) as ONE conv with doubled outputs ----
    # The recurrence tolerates TF32-rounded ACTIVATIONS (fresh rounding noise every iteration) but not TF32
    # WEIGHTS (the same perturbation 32 times): csrc/gru_fused.cu.
    @staticmethod
    def _rn_tf32(t):
        return ((t.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)

    def _split_weights(self, tag, channels_last=False, half=False):
        """``[w_hi ; w_lo]`` along the INPUT-channel axis for the (z|r) and q convolutions of one half-step (the
        activations are presented twice, ``[RN(x) ; RN(x)]``).  ``half``: both parts as fp16 (same 10-bit mantissa
        as TF32; ``w_lo`` keeps ~5 more bits before fp16's subnormal spacing cuts it, i.e. weights good to 2**-17)."""
        cz, cr, cq = (getattr(self, f"conv{g}{tag}") for g in "zrq")
        # in-place weight updates (load_state_dict, .to()) bump ``_version``: a stale split is never served
        stamp = tuple(p._version for c in (cz, cr, cq) for p in (c.weight, c.bias)) + (cz.weight.device,)
        key = (tag, channels_last, half)
        if key not in self._split_w or self._split_w[key][3] != stamp:
            packed = []
            for w, b in ((torch.cat([cz.weight, cr.weight], 0), torch.cat([cz.bias, cr.bias], 0)), (cq.weight, cq.bias)):
                w = w.detach().float()
                if half:
                    hi16 = w.clamp(-65504.0, 65504.0).half()
                    w2 = torch.cat([hi16, (w - hi16.float()).half()], 1).contiguous()
                else:
                    hi = self._rn_tf32(w)
                    w2 = torch.cat([hi, w - hi], 1).contiguous()
                if channels_last:
                    w2 = w2.contiguous(memory_format=torch.channels_last)
                packed.append((w2, b.detach().float().contiguous()))
            self._split_w[key] = (packed[0], packed[1], cz.padding, stamp)
        return self._split_w[key][:3]

    def _half_step_wsplit(self, h, x, tag):
        """The same recurrence through torch ops (the fused runner's reference and its fallback)."""
        (wzr, bzr), (wq, bq), pad = self._split_weights(tag)
        hx = self._rn_tf32(torch.cat([h, x], dim=1))
        z, r = torch.sigmoid(F.conv2d(torch.cat([hx, hx], 1), wzr, bzr, padding=pad)).chunk(2, dim=1)
        rhx = self._rn_tf32(torch.cat([r * h, x], dim=1))
        q = torch.tanh(F.conv2d(torch.cat([rhx, rhx], 1), wq, bq, padding=pad))
        return (1 - z) * h + z * q

    def _half_step_wsplit16(self, h, x, tag):
        """The fp16 form through torch ops: fp16 operands, fp32 accumulation, fp16 pre-activations (what cuDNN's
        fp16 convolution returns), fp32 gates."""
        (wzr, bzr), (wq, bq), pad = self._split_weights(tag, half=True)
        hx = torch.cat([h, x], dim=1).clamp(-65504.0, 65504.0).half()
        z, r = torch.sigmoid(F.conv2d(torch.cat([hx, hx], 1), wzr, None, padding=pad).float() + bzr.view(1, -1, 1, 1)).chunk(2, dim=1)
        rhx = torch.cat([r * h, x], dim=1).clamp(-65504.0, 65504.0).half()
        q = torch.tanh(F.conv2d(torch.cat([rhx, rhx], 1), wq, None, padding=pad).float() + bq.view(1, -1, 1, 1))
        return (1 - z) * h + z * q

    # ---- fused channels-last runner of the weight-split recurrence (csrc/gru_fused.cu) -------------------------
    def start(self, h0, inp):
        """Begin a forward: returns the per-forward runner holding the hidden state and the staging buffer."""
        return FusedGRURun(self, h0, inp, half=self.recurrence == "wsplit16")

    def forward(self, h, x):
        if self.recurrence == "wsplit" and h.is_cuda:
            with cudnn_tf32(True):
                return self._half_step_wsplit(self._half_step_wsplit(h, x, "1"), x, "2")
        if self.recurrence == "wsplit16" and h.is_cuda:
            return self._half_step_wsplit16(self._half_step_wsplit16(h, x, "1"), x, "2")
        if self.recurrence == "fp32":
            with cudnn_tf32(False):
                return self._half_step(self._half_step(h, x, "1"), x, "2")
        return self._half_step(self._half_step(h, x, "1"), x, "2")


class FusedGRURun:
    """One forward's worth of the weight-split TF32 ConvGRU on the fused channels-last kernels.

    Holds the hidden state ``h`` (channels-last) and the staging buffer ``S`` whose rows are
    ``[RN(h) | RN(x) | RN(h) | RN(x)]``; ``S`` is the NHWC input of all four convolutions of an iteration (weights
    ``[w_hi ; w_lo]`` along the input channels), so cuDNN's tensor-core kernels run without layout conversions
    and every elementwise step between two convolutions is ONE kernel (``nnd_gru_gate_r`` / ``nnd_gru_gate_h``).  The context half of ``x`` is staged once
    per forward, the motion half once per iteration.
    """

    def __init__(self, gru, h0, inp, half=False):
        from . import _lib
        self._lib = _lib
        self.gru = gru
        self.half = bool(half)      # fp16 staging buffer / weights / pre-activations (csrc/gru_fused.cu, fp16 variant)
        self.fp16_heads = bool(getattr(gru, "fp16_heads", True))
        h0, inp = h0.float(), inp.float()
        N, ch, H, W = h0.shape
        self.N, self.ch, self.H, self.W = N, ch, H, W
        self.c_inp = inp.shape[1]
        self.cx = gru.convz1.weight.shape[1] - ch
        self.ctot = 2 * (ch + self.cx)
        cl = torch.channels_last
        self.S = torch.empty(N, self.ctot, H, W, dtype=torch.float16 if self.half else torch.float32,
                             device=h0.device).contiguous(memory_format=cl)
        self.h = torch.empty(N, ch, H, W, dtype=torch.float32, device=h0.device).contiguous(memory_format=cl)
        self.z = torch.empty_like(self.h)
        # dense channels-last fp16 copy of the hidden state for the heads (fp16 form only), refreshed by every step
        self.h16 = (torch.empty(N, ch, H, W, dtype=torch.float16, device=h0.device).contiguous(memory_format=cl)
                    if self.half and self.fp16_heads else None)
        self.h.copy_(h0)
        self._stage(h0, 0)
        self._stage(inp, ch)

    def _stage(self, src, off):
        """Write ``RN_tf32(src)`` at channel offset ``off`` of the staging rows; a channels-last ``src`` is read as it lies."""
        lib = self._lib
        N, C = src.shape[0], src.shape[1]
        if not (self.half and src.dtype == torch.float16):
            src = src.float()
        cl = (not src.is_contiguous()) and src.is_contiguous(memory_format=torch.channels_last) and C % 4 == 0
        if not cl:
            src = src.float().contiguous()
        with torch.cuda.device(src.device):
            if self.half:
                kind = 0 if not cl else (2 if src.dtype == torch.float16 else 1)
                lib.check(lib.load().nnd_gru_stage_f16(src.data_ptr(), kind, N, C, self.H * self.W, self.S.data_ptr(), self.ctot,
                                                       off, lib.stream_ptr(src)), "nnd_gru_stage_f16")
                return
            lib.check(lib.load().nnd_gru_stage(lib.ptr(src), 1 if cl else 0, N, C, self.H * self.W, lib.ptr(self.S), self.ctot,
                                               off, lib.stream_ptr(src)), "nnd_gru_stage")

    def step(self, motion, flow=None):
        """One GRU update with ``x = cat[inp, motion]``; returns the new hidden state (channels-last view).

        With ``flow`` given, the last ``flow.shape[1]`` channels of ``motion`` are placeholders and the flow is
        staged over them (``motion = cat[features, flow]`` of the reference, update_block.py:65, without the cat)."""
        lib = self._lib
        if motion.shape[1] != self.cx - self.c_inp:
            raise RuntimeError(f"motion features must have {self.cx - self.c_inp} channels, got {motion.shape[1]}")
        self._stage(motion, self.ch + self.c_inp)
        if flow is not None:
            self._stage(flow, self.ch + self.cx - flow.shape[1])
        pixels = self.N * self.H * self.W
        cl = torch.channels_last
        if self.half:
            return self._step_f16(pixels)
        with cudnn_tf32(True), torch.cuda.device(self.S.device):
            for tag in "12":
                (wzr, bzr), (wq, bq), pad = self.gru._split_weights(tag, channels_last=True)
                zr = F.conv2d(self.S, wzr, None, padding=pad).contiguous(memory_format=cl)
                lib.check(lib.load().nnd_gru_gate_r(lib.ptr(zr), lib.ptr(bzr), lib.ptr(self.h), pixels, self.ch, lib.ptr(self.z),
                                                    lib.ptr(self.S), self.ctot, lib.stream_ptr(zr)), "nnd_gru_gate_r")
                q = F.conv2d(self.S, wq, None, padding=pad).contiguous(memory_format=cl)
                lib.check(lib.load().nnd_gru_gate_h(lib.ptr(q), lib.ptr(bq), lib.ptr(self.z), pixels, self.ch, lib.ptr(self.h),
                                                    lib.ptr(self.S), self.ctot, lib.stream_ptr(q)), "nnd_gru_gate_h")
        return self.h

    def _step_f16(self, pixels):
        lib = self._lib
        cl = torch.channels_last
        st = lib.stream_ptr(self.S)
        with torch.cuda.device(self.S.device):
            for tag in "12":
                (wzr, bzr), (wq, bq), pad = self.gru._split_weights(tag, channels_last=True, half=True)
                zr = F.conv2d(self.S, wzr, None, padding=pad).contiguous(memory_format=cl)
                lib.check(lib.load().nnd_gru_gate_r_f16(zr.data_ptr(), lib.ptr(bzr), lib.ptr(self.h), pixels, self.ch,
                                                        lib.ptr(self.z), self.S.data_ptr(), self.ctot, st), "nnd_gru_gate_r_f16")
                q = F.conv2d(self.S, wq, None, padding=pad).contiguous(memory_format=cl)
                h16 = self.h16.data_ptr() if (self.h16 is not None and tag == "2") else None
                lib.check(lib.load().nnd_gru_gate_h_f16(q.data_ptr(), lib.ptr(bq), lib.ptr(self.z), pixels, self.ch,
                                                        lib.ptr(self.h), self.S.data_ptr(), self.ctot, h16, st),
                          "nnd_gru_gate_h_f16")
        return self.h


class BasicMotionEncoder(nn.Module):
    """Consumer of the lookup output (reference blocks/update_block.py:39-65)."""

    def __init__(self, cor_planes, hidden_dim=128, flow_channel=2):
        super().__init__()
        self.convc1 = nn.Conv2d(cor_planes, 256, 1)
        self.convc2 = nn.Conv2d(256, 192, 3, padding=1)
        self.convf1 = nn.Conv2d(flow_channel, 128, 7, padding=3)
        self.convf2 = nn.Conv2d(128, 64, 3, padding=1)
        self.conv = nn.Conv2d(64 + 192, hidden_dim - flow_channel, 3, padding=1)
        self.channels_last = False      # set by the engine together with channels-last weights

    def forward(self, flow, corr, cor1=None, split_flow=False, half=False):
        """``cor1``: ``relu(convc1(corr))`` already computed by the fused lookup kernel (then ``corr`` is unused).
        ``half`` (with ``split_flow``): the flow branch and the last convolution run as fp16 convolutions and the
        motion features come back fp16 (they are rounded to fp16 for the recurrence anyway); ``convc2`` follows the
        dtype of ``cor1``."""
        if cor1 is not None and self.channels_last:
            # the rest of the encoder then stays channels-last: cuDNN's tensor-core kernels need no layout conversion
            cor1 = cor1.contiguous(memory_format=torch.channels_last)
        if cor1 is not None and cor1.dtype == torch.float16:
            cor = conv_relu_f16(self.convc2, cor1)      # the fused lookup kernel wrote fp16 for the fp16 iteration
        else:
            cor = conv_relu(self.convc2, cor1 if cor1 is not None else conv_relu(self.convc1, corr))
        if half and split_flow and self.channels_last:
            flo = conv_relu_f16(self.convf2, flow_conv7x7_relu(self.convf1, flow, half=True))
            if getattr(self.conv, "exact16", False):
                return conv_add_relu_split16(nhwc_cat_f16(cor, flo), self._padded_conv_split16(), self.conv.stride,
                                             self.conv.padding, self.conv.dilation), flow
            dither = getattr(self.conv, "dither16", None)
            if dither:
                extra = self.convf1.weight.shape[1]
                conv = self.conv
                w, b = half_dither_conv_params(
                    conv, *dither, tag="_f16_dither_padded",
                    weight=torch.cat([conv.weight.detach(), conv.weight.new_zeros(extra, *conv.weight.shape[1:])], 0),
                    bias=torch.cat([conv.bias.detach(), conv.bias.new_zeros(extra)], 0))
            else:
                w, b = self._padded_conv(half=True)
            return torch.cudnn_convolution_relu(nhwc_cat_f16(cor, flo), w, b, self.conv.stride, self.conv.padding,
                                                self.conv.dilation, 1), flow
        flo = conv_relu(self.convf2, flow_conv7x7_relu(self.convf1, flow) if self.channels_last
                        else conv_relu(self.convf1, flow))
        x = torch.cat([cor, flo], dim=1)
        if split_flow:
            # (motion features with one zero channel appended, flow): the caller writes the flow into that channel
            # of its staging buffer itself, so neither cuDNN's padding of the odd channel count (127) nor the
            # concatenation pass over the result is needed.  Same arithmetic: the extra filter is all zeros.
            w, b = self._padded_conv()
            return torch.cudnn_convolution_relu(x, w, b, self.conv.stride, self.conv.padding, self.conv.dilation, 1), flow
        out = conv_relu(self.conv, x)
        return torch.cat([out, flow], dim=1)

    def _padded_conv_split16(self):
        """``_padded_conv`` as the two-term fp16 split ``(w_hi, w_lo, b_hi, b_lo)`` (cached per weight version)."""
        conv = self.conv
        key = (conv.weight._version, conv.bias._version, conv.weight.device)
        cache = getattr(self, "_pad_cache16s", None)
        if cache is None or cache[0] != key:
            extra = self.convf1.weight.shape[1]
            w = torch.cat([conv.weight.detach(), conv.weight.new_zeros(extra, *conv.weight.shape[1:])], 0)
            b = torch.cat([conv.bias.detach(), conv.bias.new_zeros(extra)], 0)
            cache = (key, _split16(w, b))
            self._pad_cache16s = cache
        return cache[1]

    def _padded_conv(self, half=False):
        """``self.conv`` with zero filters appended up to ``hidden_dim`` output channels (cached per weight version)."""
        conv = self.conv
        key = (conv.weight._version, conv.bias._version, conv.weight.device, torch.backends.cudnn.allow_tf32)
        if half:
            cache = getattr(self, "_pad_cache16", None)
            if cache is None or cache[0] != key:
                w, b = half_conv_params(conv)
                extra = self.convf1.weight.shape[1]
                w = torch.cat([w, w.new_zeros(extra, *w.shape[1:])], 0).contiguous(memory_format=torch.channels_last)
                cache = (key, (w, torch.cat([b, b.new_zeros(extra)], 0).contiguous()))
                self._pad_cache16 = cache
            return cache[1]
        cache = getattr(self, "_pad_cache", None)
        if cache is None or cache[0] != key:
            w = inference_weight(conv).detach()
            extra = self.convf1.weight.shape[1]          # the flow channels the reference concatenates
            w = torch.cat([w, w.new_zeros(extra, *w.shape[1:])], 0).contiguous(memory_format=torch.channels_last)
            b = torch.cat([conv.bias.detach(), conv.bias.new_zeros(extra)], 0).contiguous()
            cache = (key, (w, b))
            self._pad_cache = cache
        return cache[1]


class FlowHead(nn.Module):
    def __init__(self, input_dim=128, hidden_dim=256, flow_channel=2):
        super().__init__()
        self.conv1 = nn.Conv2d(input_dim, hidden_dim, 3, padding=1)
        self.conv2 = nn.Conv2d(hidden_dim, flow_channel, 3, padding=1)
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x, coords=None, org=None):
        """``delta``; with ``coords`` / ``org`` and the fused tail available: ``(coords + delta, coords + delta - org)``."""
        hidden = conv_relu_f16(self.conv1, x) if x.dtype == torch.float16 else conv_relu(self.conv1, x)
        fused = flow_head_tail(self.conv2, hidden, coords, org)
        if fused is not None:
            return fused
        delta = conv_plain(self.conv2, hidden.float())
        return delta if coords is None else (coords + delta, coords + delta - org)


class BasicUpdateBlock(nn.Module):
    """Motion encoder -> GRU -> flow head + 0.25 * mask head (reference blocks/update_block.py:68-112)."""

    def __init__(self, hidden_dim, cor_planes, context_dim=128, gru="sep_conv", flow_channel=2, spatial_scale=8):
        super().__init__()
        if gru != "sep_conv":
            raise NotImplementedError("only the separable ConvGRU of BaseRAFTStereo is provided")
        self.encoder = BasicMotionEncoder(cor_planes, hidden_dim=hidden_dim, flow_channel=flow_channel)
        self.gru = SepConvGRU(hidden_dim=hidden_dim, input_dim=context_dim + hidden_dim)
        self.flow_head = FlowHead(hidden_dim, hidden_dim=hidden_dim, flow_channel=flow_channel)
        sps = spatial_scale ** 2 if isinstance(spatial_scale, int) else spatial_scale[0] * spatial_scale[1]
        self.mask = nn.Sequential(nn.Conv2d(hidden_dim, hidden_dim * 2, 3, padding=1), nn.ReLU(inplace=True),
                                  nn.Conv2d(hidden_dim * 2, sps * 9, 1))

    def forward(self, net, inp, corr, flow, raw_mask=False, cor1=None, gru_run=None, coords=None, org=None, need_mask=True):
        """``raw_mask=True`` returns the mask logits without the last convolution's bias and without the reference's
        ``0.25 *`` (update_block.py:110): the fused upsampling kernel applies both, saving two passes over the
        (N,576,H,W) tensor.  With ``coords`` / ``org`` the third result is ``(coords + delta, coords + delta - org)``
        (the loop's update, fused into the flow head's last convolution) instead of ``delta``.  ``need_mask=False``
        (``final_only`` forwards, every iteration but the last) skips the mask head: the mask only feeds the upsampling
        of that iteration's prediction, never the recurrence."""
        if gru_run is not None:
            split = flow.is_cuda and not torch.is_grad_enabled() and hasattr(torch, "cudnn_convolution_relu")
            motion = self.encoder(flow, corr, cor1=cor1, split_flow=split, half=split and getattr(gru_run, "half", False))
            # fused channels-last weight-split recurrence; `net` lives in the runner
            net = gru_run.step(*motion) if split else gru_run.step(motion)
        else:
            motion = self.encoder(flow, corr, cor1=cor1)
            net = self.gru(net, torch.cat((inp, motion), dim=1))
        net16 = getattr(gru_run, "h16", None) if (gru_run is not None and raw_mask and coords is not None) else None
        if not need_mask:
            return net, None, self.flow_head(net16 if net16 is not None else net, coords, org)
        if net16 is not None:
            # fp16 recurrence: the heads read the fp16 copy of the new hidden state and run as fp16 convolutions too
            # (same operand mantissa as TF32, twice the rate); the logits stay fp16 for the fused upsampling kernel
            hidden = conv_relu_f16(self.mask[0], net16)
            mask = F.conv2d(hidden, half_conv_params(self.mask[2])[0], None, self.mask[2].stride, self.mask[2].padding)
            return net, mask, self.flow_head(net16, coords, org)
        hidden = conv_relu(self.mask[0], net)
        if raw_mask:
            # bias-free logits: the fused upsampling kernel adds self.mask[2].bias and applies the 0.25 itself
            last = self.mask[2]
            mask = F.conv2d(hidden, inference_weight(last), None, last.stride, last.padding)
        else:
            mask = 0.25 * conv_plain(self.mask[2], hidden)
        return net, mask, self.flow_head(net, coords, org)


def convex_upsample(flow, mask, rate=8):
    """Convex combination of the 3x3 coarse neighbourhood (reference raft_stereo/model.py:93-105)."""
    N, _, H, W = flow.shape
    mask = torch.softmax(mask.view(N, 1, 9, rate, rate, H, W), dim=2)
    up = F.unfold(rate * flow, (3, 3), padding=1).view(N, 1, 9, 1, 1, H, W)
    up = torch.sum(mask * up, dim=2).permute(0, 1, 4, 2, 5, 3)
    return up.reshape(N, 1, rate * H, rate * W)


class RAFTStereo(nn.Module):
    """Reference ``RAFTStereo`` API (model.py:17-139) with the correlation path on B200 kernels."""

    def __init__(self, iters=12, fnet_dim=256, hidden_dim=128, context_dim=128, corr_levels=4, corr_radius=4,
                 tracing=False, include_preprocessing=False, weights=None, strict_load=True, **kwargs):
        super().__init__()
        self.iters = iters
        self.fnet_dim = fnet_dim
        self.hidden_dim = hidden_dim
        self.context_dim = context_dim
        self.corr_levels = corr_levels
        self.corr_radius = corr_radius
        self.fnet = self._init_fnet(**kwargs)
        self.cnet_proj = nn.Sequential(nn.Conv2d(fnet_dim, context_dim + hidden_dim, 3, padding=1), nn.ReLU(False))
        self.update_block = self._init_update_block()
        self.corr_fn = CorrBlock1D
        self.tracing = tracing
        self.include_preprocessing = include_preprocessing
        self.weights = weights
        self.strict_load = strict_load
        self.final_only = False     # True: upsample only the last iteration (what evaluate.py:155 consumes)
        self.fuse_motion_front = True   # lookup + convc1 + ReLU as one kernel when corr_fn provides it
        self.fuse_gru = True            # weight-split ConvGRU on the fused channels-last kernels (mode "mixed2x")
        # Precision of the dense (cuDNN) layers, measured on the KITTI/32-iteration golden (profiles/r2_exp_modules.jsonl):
        #   "fp32"  every convolution in fp32                           final EPE vs reference 0.0002 px
        #   "mixed" ConvGRU in fp32, everything else on TF32 tensor cores              0.0021 px  (bar: 0.01 px)
        #   "mixed2x" as "mixed", the ConvGRU with TF32 activations x split fp32 weights [w_hi; w_lo] on tensor
        #           cores (the recurrence is sensitive to weight rounding only)          0.0031 px, 3x faster
        #   "tf32"  everything TF32 (PyTorch's CUDA default)                           0.0147 px  -> outside the bar
        # None leaves torch.backends.cudnn.allow_tf32 as the caller set it.
        self.dense_precision = None
        # mixed16 only.  exact_weights: the convolutions INSIDE the refinement loop take their weights as the two-term
        # fp16 split w_hi + w_lo (a rounded weight is the same perturbation in all 32 iterations and accumulates
        # coherently; gpurun_out/r2_exp_modules.log: with weight seeds 1 / 2 on the shipped KITTI pair single-term
        # weights put the final disparity 0.013 - 0.018 px from the fp32 reference, outside the 0.01 px bar).
        # exact_encoder: the same for the feature encoder (run once per forward).
        self.exact_weights = True
        self.exact_encoder = True
        # dither_weights = K > 0 (instead of exact_weights): the loop's convolutions take ONE fp16 product with variant
        # (iteration mod K) of K dithered roundings of their weights -- see _dither16
        self.dither_weights = 0
        self._graphs = {}
        if weights is not None:
            state = torch.load(weights, map_location="cpu") if not str(weights).endswith(".safetensors") else None
            if state is None:
                from safetensors.torch import load_file
                state = load_file(weights)
            self.load_state_dict(state.get("model", state), strict=strict_load)

    def _init_fnet(self):
        raise NotImplementedError("Must be implemented in child class")

    def _init_update_block(self):
        raise NotImplementedError("Must be implemented in child class")

    def forward_fnet(self, frame1, frame2):
        raise NotImplementedError("Must be implemented in child class")

    def freeze_bn(self):
        for m in self.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.eval()

    def initialize_coords(self, fmap1):
        B, _, H, W = fmap1.shape
        return torch.arange(W, device=fmap1.device).float()[None, None, None, :].repeat(B, 1, H, 1)

    convex_upsample = staticmethod(convex_upsample)

    def forward(self, frame1, frame2, **kwargs):
        if self.dense_precision is None:
            return self._forward(frame1, frame2, **kwargs)
        if self.dense_precision not in ("fp32", "mixed", "mixed2x", "mixed16", "tf32"):
            raise ValueError("dense_precision must be None, 'fp32', 'mixed', 'mixed2x', 'mixed16' or 'tf32', "
                             f"got {self.dense_precision!r}")
        # the mode is applied for the duration of this call only: the submodule switches it drives are restored on
        # the way out, so a later call with ``dense_precision = None`` really runs on the caller's own flags
        gru = getattr(self.update_block, "gru", None)
        has_half = hasattr(self.fnet, "half_convs") or isinstance(self.fnet, BasicEncoder)
        saved = (getattr(gru, "recurrence", None), getattr(self.fnet, "half_convs", False))
        if gru is not None:
            gru.recurrence = {"mixed": "fp32", "mixed2x": "wsplit", "mixed16": "wsplit16"}.get(self.dense_precision)
        if has_half:
            # mixed16 runs the (BatchNorm-folded) feature encoder as fp16 convolutions as well: the feature maps are
            # rounded to 10 mantissa bits for the correlation volume anyway (RN_tf32(RN_fp16(x)) == RN_fp16(x))
            self.fnet.half_convs = self.dense_precision == "mixed16" and getattr(self, "fp16_encoder", True)
        self._set_exact16(self.dense_precision == "mixed16" and self.exact_weights and not self.dither_weights,
                          self.dense_precision == "mixed16" and self.exact_encoder)
        try:
            with cudnn_tf32(self.dense_precision != "fp32"):
                return self._forward(frame1, frame2, **kwargs)
        finally:
            self._set_exact16(False, False)
            if gru is not None:
                gru.recurrence = saved[0]
            if has_half:
                self.fnet.half_convs = saved[1]

    def _loop_convs(self):
        ub = self.update_block
        enc, head = getattr(ub, "encoder", None), getattr(ub, "flow_head", None)
        return [c for c in (getattr(enc, "convc2", None), getattr(enc, "convf2", None), getattr(enc, "conv", None),
                            getattr(head, "conv1", None)) if c is not None]

    def _set_exact16(self, iteration, encoder):
        """Switch the two-term fp16 weights (``w_hi + w_lo``, see ``conv_add_relu_split16``) of the convolutions inside
        the refinement loop (motion encoder, flow head; the ConvGRU has its own split) and of the feature encoder."""
        for conv in self._loop_convs():
            conv.exact16 = bool(iteration)
            conv.dither16 = None
        for m in self.fnet.modules():
            if isinstance(m, (BasicEncoder, ResidualBlock)):
                m.exact16 = bool(encoder)
        self.cnet_proj[0].exact16 = bool(encoder)

    def _forward(self, frame1, frame2, **kwargs):
        fmap1, fmap2, cnet1 = self.forward_fnet(frame1, frame2)
        fnet_ds = frame1.shape[-1] // fmap1.shape[-1]
        from . import corr as _corr
        if not (isinstance(self.corr_fn, type) and issubclass(self.corr_fn, _corr.CorrBlock1D)
                and not torch.is_grad_enabled() and _corr._half_channels_last(fmap1, fmap2, None)):
            # (an fp16 channels-last pair from the fp16 encoder goes to CorrBlock1D as it is: the build reads it in place)
            fmap1, fmap2 = fmap1.float(), fmap2.float()
        net, inp = torch.split(cnet1, [self.hidden_dim, self.context_dim], dim=1)
        net, inp = torch.tanh(net.float()), F.relu(inp)

        corr = self.corr_fn(fmap1, fmap2, self.corr_levels, self.corr_radius)
        gru = getattr(self.update_block, "gru", None)
        gru_run = None
        if (self.fuse_gru and gru is not None and getattr(gru, "recurrence", None) in ("wsplit", "wsplit16") and net.is_cuda
                and not torch.is_grad_enabled()):
            gru_run = gru.start(net, inp)
        org_coords = self.initialize_coords(fmap1)
        coords1 = org_coords.clone()
        outputs = []
        dither = self.dither_weights if (self.dense_precision == "mixed16" and not torch.is_grad_enabled()) else 0
        for it in range(self.iters):
            coords1 = coords1.detach()
            if dither:
                for conv in self._loop_convs():
                    conv.dither16 = (it % dither, dither)
            fuse_front = (self.fuse_motion_front and coords1.is_cuda and not torch.is_grad_enabled()
                          and hasattr(corr, "lookup_conv1x1") and self.corr_levels == 4 and self.corr_radius == 4)
            if fuse_front:
                # lookup + convc1 (1x1) + ReLU in one kernel: the motion encoder's input is written directly
                conv1 = self.update_block.encoder.convc1
                if it == 0:
                    conv1_wt = corr.prepare_conv1x1_weight(conv1.weight)    # k-major copy, once per forward
                tf32 = bool(torch.backends.cudnn.allow_tf32)
                cl_out = tf32 and self.update_block.encoder.channels_last
                sampled, cor1 = None, corr.lookup_conv1x1(coords1, None, conv1.bias, relu=True, weight_t=conv1_wt,
                                                          precision="tf32" if tf32 else "fp32", channels_last=cl_out,
                                                          half=cl_out and getattr(gru_run, "half", False))
            else:
                sampled, cor1 = corr(coords1), None
            # on the GPU the convex upsampling is one fused kernel (softmax + unfold + weighted sum + pixel
            # shuffle, with the update block's 0.25 mask scale folded in); the torch chain below is the
            # reference's own, kept for the CPU baseline leg (corr_fn = oracle) only
            fused = coords1.is_cuda and fnet_ds in (2, 4, 8) and not torch.is_grad_enabled()
            if it == 0:
                flow = coords1 - org_coords
            want_up = not self.final_only or it == self.iters - 1
            if fused:
                # the flow head's last convolution also writes the new coordinates and the new flow
                net, mask, (coords1, flow) = self.update_block(net, inp, sampled, flow, raw_mask=True, cor1=cor1,
                                                               gru_run=gru_run, coords=coords1, org=org_coords,
                                                               need_mask=want_up)
            else:
                net, mask, delta = self.update_block(net, inp, sampled, flow, raw_mask=False, cor1=cor1, gru_run=gru_run,
                                                     need_mask=want_up)
                coords1 = coords1 + delta
                flow = coords1 - org_coords
            if want_up:
                if fused:
                    up = fused_convex_upsample(flow, mask, rate=fnet_ds, mask_scale=0.25,
                                               mask_bias=self.update_block.mask[2].bias)
                else:
                    up = self.convex_upsample(flow, mask, rate=fnet_ds)
                outputs.append({"up_disp": up})
        return outputs

    # ---- CUDA-graph replay of the whole forward (inference) ---------------------------------------
    def _graph_key(self, frame1):
        """Everything a captured graph bakes in.  Python does not run on replay, and the derived weights built during
        warm-up (TF32-rounded, BatchNorm-folded, split, padded copies) are captured by ADDRESS: the key therefore
        carries a stamp of every parameter's / buffer's in-place version and storage, next to the shape and every
        switch that changes which kernels run.  ``load_state_dict`` / ``.to()`` / ``.half()`` also drop the cache
        outright (hooks below) so stale graphs do not pile up."""
        from . import corr as _corr
        stamp = 0
        for t in itertools.chain(self.parameters(), self.buffers()):
            stamp = (stamp * 1000003 + t._version * 31 + (t.data_ptr() >> 4)) & 0xFFFFFFFFFFFF
        gru = getattr(self.update_block, "gru", None)
        enc = getattr(self.update_block, "encoder", None)
        return (tuple(frame1.shape), tuple(frame1.stride()), frame1.device.index, frame1.dtype, self.iters, self.final_only,
                self.dense_precision, stamp, self.corr_fn, _corr.get_volume_precision(), self.fuse_motion_front,
                self.fuse_gru, getattr(self, "fp16_encoder", True), self.exact_weights, self.exact_encoder, self.dither_weights,
                getattr(gru, "recurrence", None),
                getattr(gru, "_fuse_zr", False), getattr(self.fnet, "half_convs", False),
                getattr(enc, "channels_last", False), self.corr_levels, self.corr_radius,
                bool(torch.backends.cudnn.allow_tf32) if self.dense_precision is None else None,
                bool(torch.backends.cudnn.benchmark))

    def _apply(self, fn, *args, **kwargs):
        self._graphs = {}           # .to() / .cuda() / .half(): every captured address is stale
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._graphs = {}
        return super().load_state_dict(*args, **kwargs)

    @torch.no_grad()
    def forward_graphed(self, frame1, frame2):
        """Replay ``forward`` as one CUDA graph per input shape; returns the static output list.

        Every kernel of the path (cuDNN convs and this package's launches) is stream-ordered with no
        host synchronisation, so the 32-iteration loop captures cleanly.  Outputs are overwritten by
        the next replay of the same shape.
        """
        key = self._graph_key(frame1)
        entry = self._graphs.get(key)
        if entry is None:
            static1, static2 = torch.empty_like(frame1), torch.empty_like(frame2)
            static1.copy_(frame1)
            static2.copy_(frame2)
            side = torch.cuda.Stream(device=frame1.device)
            side.wait_stream(torch.cuda.current_stream(frame1.device))
            with torch.cuda.stream(side):
                for _ in range(2):      # warm-up outside capture: cuDNN autotune, lazy module init
                    self.forward(static1, static2)
            torch.cuda.current_stream(frame1.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = self.forward(static1, static2)
            entry = (graph, static1, static2, static_out)
            self._graphs[key] = entry
        graph, static1, static2, static_out = entry
        static1.copy_(frame1, non_blocking=True)
        static2.copy_(frame2, non_blocking=True)
        graph.replay()
        return static_out


class BaseRAFTStereo(RAFTStereo):
    """The paper's configuration (reference model.py:142-163)."""

    def _init_fnet(self):
        return BasicEncoder(output_dim=self.fnet_dim)

    def _init_update_block(self):
        return BasicUpdateBlock(hidden_dim=self.hidden_dim, cor_planes=self.corr_levels * (self.corr_radius * 2 + 1),
                                flow_channel=1, context_dim=self.context_dim, spatial_scale=8)

    def forward_fnet(self, frame1, frame2):
        fmap1, fmap2 = self.fnet([frame1, frame2])
        cnet = conv_relu_f16(self.cnet_proj[0], fmap1) if fmap1.dtype == torch.float16 else conv_relu(self.cnet_proj[0], fmap1)
        return fmap1, fmap2, cnet

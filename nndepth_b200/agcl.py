"""CREStereo adaptive group correlation layer on the sm_100a kernels.

Mirrors ``nndepth/models/cre_stereo/cost_volume.py:6-154`` of the reference (class ``AGCL``): same
constructor, ``__call__(flow, extra_offset, small_patch=False, iter_mode=False)`` and method names.
The reference model names ``AGCL`` directly (``cre_stereo/model.py:198-200``), so it is swapped by
patching that module attribute.  The optional LoFTR cross-attention (``att``) is dense attention and
stays the caller's PyTorch module; it is a pure function of the two maps, so its output is computed
once and cached instead of once per call (reference :91-99).

Layout: the kernels gather whole channel vectors, so for ``C % 16 == 0`` (every model configuration) the
two maps are staged ONCE per object as channels-last ``(N, H, W, C)`` copies (``nnd_nchw_to_nhwc``) and all
6-12 calls of a cascade scale run on those; other channel counts take the generic NCHW kernels.
"""
import torch

from . import _lib
from .corr import _train_f32, _wants_grad


class _AGCLOffset(torch.autograd.Function):
    """Differentiable offset mode on channels-last maps: forward ``nnd_agcl_offset_nhwc``, backward
    ``nnd_agcl_offset_backward_nhwc`` (gradients for both maps, the flow and the learned offsets)."""

    @staticmethod
    def forward(ctx, left, right, flow, extra, small_patch):
        ctx.save_for_backward(left, right, flow, extra)
        ctx.small = bool(small_patch)
        return _lib.ops().agcl_offset(left, right, flow, extra, ctx.small, True)

    @staticmethod
    def backward(ctx, grad_out):
        left, right, flow, extra = ctx.saved_tensors
        d1, d2, dflow, dextra = _lib.ops().agcl_offset_backward(left, right, flow, extra, grad_out.contiguous().float(), ctx.small)
        need = ctx.needs_input_grad
        return (d1 if need[0] else None, d2 if need[1] else None, dflow if need[2] else None, dextra if need[3] else None, None)


class _AGCLIter(torch.autograd.Function):
    """Differentiable iter mode on channels-last maps: forward ``nnd_agcl_iter_nhwc`` (one fused pass); backward
    re-materialises the flow-warped right map (``nnd_agcl_warp_nhwc``) and runs ``nnd_agcl_iter_backward_nhwc``."""

    @staticmethod
    def forward(ctx, left, right, flow, small_patch, detach_warped):
        ctx.save_for_backward(left, right, flow)
        ctx.small, ctx.left_only = bool(small_patch), bool(detach_warped)
        ws = torch.empty_like(right)
        return _lib.ops().agcl_iter(left, right, flow, ctx.small, True, ws)

    @staticmethod
    def backward(ctx, grad_out):
        left, right, flow = ctx.saved_tensors
        warped = _lib.ops().agcl_warp(right, flow)
        d1, d2, dflow = _lib.ops().agcl_iter_backward(left, right, flow, warped, grad_out.contiguous().float(), ctx.small,
                                                      ctx.left_only)
        need = ctx.needs_input_grad
        if ctx.left_only:
            return (d1 if need[0] else None, None, None, None, None)
        return (d1 if need[0] else None, d2 if need[1] else None, dflow if need[2] else None, None, None)


class AGCL:
    def __init__(self, fmap1, fmap2, att=None):
        # training (feature maps, or the attention module's parameters, require grad): the maps stay attached to the
        # autograd graph and the calls go through _AGCLOffset / _AGCLIter; otherwise inference on detached copies
        self._train = _wants_grad(fmap1, fmap2) or (
            torch.is_grad_enabled() and isinstance(att, torch.nn.Module) and any(p.requires_grad for p in att.parameters()))
        self.fmap1 = _train_f32(fmap1, "fmap1") if self._train else _lib.as_cuda_f32(fmap1, "fmap1")
        self.fmap2 = _train_f32(fmap2, "fmap2") if self._train else _lib.as_cuda_f32(fmap2, "fmap2")
        if self.fmap1.dim() != 4 or self.fmap1.shape != self.fmap2.shape:
            raise RuntimeError("fmap1 and fmap2 must be (N, C, H, W) of identical shape")
        self.att = att
        # iter mode: the reference replicate-pads a DETACHED clone of the warped right map (manual_pad, cre_stereo/utils.py:
        # 29-31), so its gradient reaches the left features only; False differentiates through the warp as well
        self.detach_warped = True
        self._attended = None
        self._staged = {}       # id(nchw tensor) -> (nchw tensor kept alive, channels-last copy)
        self._warp_ws = None    # workspace of the flow-warped right map (iter mode)

    @staticmethod
    def _fast(C):
        return C % 16 == 0 and C <= 512

    def _nhwc(self, t):
        """Channels-last staging copy of an ``(N, C, H, W)`` map, made once per tensor."""
        hit = self._staged.get(id(t))
        if hit is not None and hit[0] is t:
            return hit[1]
        out = _lib.ops().nchw_to_nhwc(t)
        if len(self._staged) >= 4:          # fmap1, fmap2 and their attended versions; transient maps rotate out
            self._staged.pop(next(iter(self._staged)))
        self._staged[id(t)] = (t, out)
        return out

    def _training_call(self, *tensors):
        """Differentiable path: grad mode on and something upstream requires grad (maps with C % 16 == 0 only)."""
        if not (torch.is_grad_enabled() and (self._train or any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors))):
            return False
        C = tensors[0].shape[1]
        if C % 16 != 0:
            raise RuntimeError(f"the differentiable AGCL path needs C % 16 == 0 (got {C})")
        return True

    @staticmethod
    def _nhwc_grad(t):
        """Channels-last copy that stays in the autograd graph (the staging kernel of the inference path does not)."""
        return t.float().permute(0, 2, 3, 1).contiguous()

    @staticmethod
    def _flow_grad(flow, N, H, W):
        if not (isinstance(flow, torch.Tensor) and flow.is_cuda):
            raise RuntimeError("flow must be a CUDA tensor: nndepth_b200 has no CPU path")
        if tuple(flow.shape) != (N, 2, H, W):
            raise RuntimeError(f"flow must be (N, 2, H, W) = {(N, 2, H, W)}, got {tuple(flow.shape)}")
        return flow.float().contiguous()

    def __call__(self, flow, extra_offset, small_patch=False, iter_mode=False):
        if iter_mode:
            return self.corr_iter(self.fmap1, self.fmap2, flow, small_patch)
        return self.corr_att_offset(self.fmap1, self.fmap2, flow, extra_offset, small_patch)

    @staticmethod
    def _check_flow(flow, N, H, W):
        flow = _lib.as_cuda_f32(flow, "flow")
        if tuple(flow.shape) != (N, 2, H, W):
            raise RuntimeError(f"flow must be (N, 2, H, W) = {(N, 2, H, W)}, got {tuple(flow.shape)}")
        return flow

    def corr_iter(self, left_feature, right_feature, flow, small_patch):
        left = left_feature if left_feature is self.fmap1 else _lib.as_cuda_f32(left_feature, "left_feature")
        right = right_feature if right_feature is self.fmap2 else _lib.as_cuda_f32(right_feature, "right_feature")
        N, C, H, W = left.shape
        if self._training_call(left, right, flow):
            return _AGCLIter.apply(self._nhwc_grad(left), self._nhwc_grad(right), self._flow_grad(flow, N, H, W), bool(small_patch),
                                   self.detach_warped)
        flow = self._check_flow(flow, N, H, W)
        if self._fast(C) and H >= 2 and W >= 2:
            if self._warp_ws is None or self._warp_ws.shape != (N, H, W, C) or self._warp_ws.device != left.device:
                self._warp_ws = torch.empty(N, H, W, C, dtype=torch.float32, device=left.device)
            return _lib.ops().agcl_iter(self._nhwc(left), self._nhwc(right), flow, bool(small_patch), True, self._warp_ws)
        return _lib.ops().agcl_iter(left, right, flow, bool(small_patch), False, None)

    def _attend(self, left, right):
        """Cross-attention on ``(N, H*W, C)`` token layout and back (reference :91-99), cached."""
        if self._attended is None or self._attended[0] is not left or self._attended[1] is not right:
            N, C, H, W = left.shape
            lt = left.permute(0, 2, 3, 1).reshape(N, H * W, C)
            rt = right.permute(0, 2, 3, 1).reshape(N, H * W, C)
            lt, rt = self.att(lt, rt)
            la, ra = [x.reshape(N, H, W, C).permute(0, 3, 1, 2) for x in (lt, rt)]
            if self._train and torch.is_grad_enabled():
                self._attended = (left, right, _train_f32(la, "att(left)"), _train_f32(ra, "att(right)"))
            else:
                self._attended = (left, right, _lib.as_cuda_f32(la, "att(left)"), _lib.as_cuda_f32(ra, "att(right)"))
        return self._attended[2], self._attended[3]

    def corr_att_offset(self, left_feature, right_feature, flow, extra_offset, small_patch):
        left = left_feature if left_feature is self.fmap1 else _lib.as_cuda_f32(left_feature, "left_feature")
        right = right_feature if right_feature is self.fmap2 else _lib.as_cuda_f32(right_feature, "right_feature")
        N, C, H, W = left.shape
        if self.att is not None:
            left, right = self._attend(left, right)
        if self._training_call(left, right, flow, extra_offset):
            extra = _train_f32(extra_offset, "extra_offset")
            if tuple(extra.shape) != (N, 18, H, W):
                raise RuntimeError(f"extra_offset must be (N, 18, H, W) = {(N, 18, H, W)}, got {tuple(extra.shape)}")
            return _AGCLOffset.apply(self._nhwc_grad(left), self._nhwc_grad(right), self._flow_grad(flow, N, H, W), extra,
                                     bool(small_patch))
        flow = self._check_flow(flow, N, H, W)
        extra = _lib.as_cuda_f32(extra_offset, "extra_offset")
        if tuple(extra.shape) != (N, 18, H, W):
            raise RuntimeError(f"extra_offset must be (N, 18, H, W) = {(N, 18, H, W)}, got {tuple(extra.shape)}")
        if self._fast(C) and H >= 2 and W >= 2:
            return _lib.ops().agcl_offset(self._nhwc(left), self._nhwc(right), flow, extra, bool(small_patch), True)
        return _lib.ops().agcl_offset(left, right, flow, extra, bool(small_patch), False)

    def get_correlation(self, left_feature, right_feature, psize=(3, 3), dilate=(1, 1)):
        """Replicate-padded local correlation of ONE channel group -> ``(N, 9, H, W)`` (reference :28-52).

        Served by the iter-mode kernel with zero flow (warping by zero flow is the identity up to the
        reference's own fp32 coordinate round trip), on a 4x channel-replicated input, group 0 returned.
        """
        if tuple(dilate) != (1, 1) or tuple(psize) not in ((3, 3), (1, 9)):
            raise NotImplementedError("only the (3,3) and (1,9) unit-dilation windows the model uses are built")
        left = _lib.as_cuda_f32(left_feature, "left_feature")
        right = _lib.as_cuda_f32(right_feature, "right_feature")
        N, C, H, W = left.shape
        zero = torch.zeros(N, 2, H, W, dtype=torch.float32, device=left.device)
        full = self.corr_iter(left.repeat(1, 4, 1, 1), right.repeat(1, 4, 1, 1), zero, tuple(psize) == (3, 3))
        return full[:, :9].contiguous()

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


class AGCL:
    def __init__(self, fmap1, fmap2, att=None):
        self.fmap1 = _lib.as_cuda_f32(fmap1, "fmap1")
        self.fmap2 = _lib.as_cuda_f32(fmap2, "fmap2")
        if self.fmap1.dim() != 4 or self.fmap1.shape != self.fmap2.shape:
            raise RuntimeError("fmap1 and fmap2 must be (N, C, H, W) of identical shape")
        self.att = att
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
            self._attended = (left, right, _lib.as_cuda_f32(la, "att(left)"), _lib.as_cuda_f32(ra, "att(right)"))
        return self._attended[2], self._attended[3]

    def corr_att_offset(self, left_feature, right_feature, flow, extra_offset, small_patch):
        left = left_feature if left_feature is self.fmap1 else _lib.as_cuda_f32(left_feature, "left_feature")
        right = right_feature if right_feature is self.fmap2 else _lib.as_cuda_f32(right_feature, "right_feature")
        N, C, H, W = left.shape
        if self.att is not None:
            left, right = self._attend(left, right)
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

"""Convex upsampling of the coarse disparity on the sm_100a kernel -- one pass instead of the reference's
softmax / unfold / multiply / sum / permute chain (``nndepth/models/raft_stereo/model.py:93-105``; the same
code serves CREStereo ``cre_stereo/model.py:110-122`` and IGEV ``igev_stereo/model.py:103-115``).
"""
import torch

from . import _lib


def convex_upsample(flow, mask, rate=8, mask_scale=1.0, mask_bias=None):
    """``flow (N,1,H,W)``, ``mask (N, 9*rate*rate, H, W)`` -> ``(N, 1, rate*H, rate*W)``.

    ``mask_scale`` multiplies the mask logits inside the kernel: ``0.25`` folds the update block's
    ``0.25 * mask`` (``blocks/update_block.py:110``) into the same pass (exact: a power of two).
    ``mask_bias`` ``(9*rate*rate,)`` is added to the logits first -- the bias of the mask head's last 1x1
    convolution, which can then run bias-free.
    """
    flow = _lib.as_cuda_f32(flow, "flow")
    # a channels-last mask (what cuDNN returns for a channels-last hidden state) is consumed as it lies
    # (fp32, or fp16 when the mask head ran as fp16 convolutions)
    nhwc = (rate == 8 and isinstance(mask, torch.Tensor) and mask.is_cuda and mask.dtype in (torch.float32, torch.float16)
            and mask.dim() == 4
            and mask.shape[1] > 1 and not mask.is_contiguous() and mask.is_contiguous(memory_format=torch.channels_last)
            and not (torch.is_grad_enabled() and mask.requires_grad))
    mask = mask.detach() if nhwc else _lib.as_cuda_f32(mask, "mask")
    if flow.dim() != 4 or flow.shape[1] != 1:
        raise RuntimeError(f"flow must be (N, 1, H, W), got {tuple(flow.shape)}")
    N, _, H, W = flow.shape
    if tuple(mask.shape) != (N, 9 * rate * rate, H, W):
        raise RuntimeError(f"mask must be (N, 9*rate*rate, H, W) = {(N, 9 * rate * rate, H, W)}, got {tuple(mask.shape)}")
    if mask_bias is not None:
        mask_bias = _lib.as_cuda_f32(mask_bias, "mask_bias")
        if mask_bias.numel() != 9 * rate * rate:
            raise RuntimeError(f"mask_bias must have {9 * rate * rate} elements, got {mask_bias.numel()}")
    return _lib.ops().convex_upsample(flow, mask, mask_bias, int(rate), float(mask_scale),
                                      (2 if mask.dtype == torch.float16 else 1) if nhwc else 0)

/*
 * nndepth_b200.h -- C ABI of the B200-native stereo-correlation hot path.
 *
 * The reference (anhtu293/nndepth) is pure Python/PyTorch and has NO plugin / FFI boundary of its
 * own; the drop-in boundary is the constructor/__call__ of its correlation classes.
 * Each entry point below therefore cites the reference *function* it replaces (file:line relative to
 * the reference root) -- the Python mirror classes in nndepth_b200/ are the only callers.
 *
 * Conventions
 *   - plain C types only: device pointers, ints, an opaque stream handle (a cudaStream_t);
 *   - every tensor is dense fp32 in the layout named per function; "pitch" = row stride in floats;
 *   - nothing here allocates device memory that outlives the call, except TMA descriptors cached
 *     per shape inside the library; outputs are caller-allocated; inputs are never written;
 *   - all work is enqueued on `stream`, no host synchronisation (CUDA-graph capturable);
 *   - return value: NND_OK or an error code; nnd_last_error_string() describes the last failure of
 *     the calling thread.  Nothing throws or aborts across this boundary.
 *   - there is no CPU fallback: on a machine without a CUDA device every compute entry point
 *     returns NND_ERR_CUDA.
 */
#ifndef NNDEPTH_B200_H
#define NNDEPTH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NND_ABI_VERSION 1
#define NND_MAX_LEVELS 8

typedef int nnd_status;
enum {
  NND_OK = 0,
  NND_ERR_INVALID_ARGUMENT = 1,
  NND_ERR_UNSUPPORTED = 2,
  NND_ERR_CUDA = 3
};

/* precision of the volume contraction (operands; accumulation is always fp32) */
enum {
  NND_PREC_FP32 = 0,  /* CUDA-core FFMA, bit-for-bit fp32 operands (parity mode, 1e-5 bar)        */
  NND_PREC_TF32 = 1   /* tcgen05 kind::tf32, operands truncated to 10-bit mantissa (1e-3 bar)     */
};

typedef struct CUstream_st* nnd_stream_t; /* == cudaStream_t */

int nnd_abi_version(void);
const char* nnd_last_error_string(void);

/* Number of floats the caller must allocate for a pyramid row at level width `w` (16-byte rows). */
int nnd_row_pitch(int width);

/* ------------------------------------------------------------------------------------------------
 * RAFT-Stereo 1-D correlation pyramid.
 * Replaces CorrBlock1D.corr + CorrBlock1D.__init__, nndepth/models/raft_stereo/cost_volume.py:55-61,
 * :12-34 (torch.matmul / C**0.5, then avg_pool1d(.,2) per level) in ONE pass: the volume is written
 * once and the pooled levels come out of the same epilogue.
 *   fmap1 (B,C,H,W1), fmap2 (B,C,H,W2) NCHW contiguous.
 *   level[l] (l < num_levels): rows = B*H*W1, width w_l = W2 >> l (floor), row pitch pitch[l] floats.
 *   num_levels in [1, NND_MAX_LEVELS]; every written level must have width >= 1.
 * ---------------------------------------------------------------------------------------------- */
nnd_status nnd_corr1d_build(const float* fmap1, const float* fmap2, int B, int C, int H, int W1, int W2,
                            int num_levels, int precision, float* const* level, const int* pitch,
                            nnd_stream_t stream);

/* nnd_corr1d_build for feature maps that are fp16 and channels-last, (B, H, W, C) in memory -- what a cuDNN fp16 encoder
 * returns (raft_stereo/model.py:111-124 hands the encoder output to CorrBlock1D).  fp16 x fp16 products with fp32
 * accumulation on tcgen05 (the products are exact, like TF32 products of the same values); C % 8 == 0, W1 % 4 == 0,
 * W2 % 4 == 0.  Same pyramid as nnd_corr1d_build(NND_PREC_TF32) on the values converted to fp32 NCHW, to fp32
 * accumulation order. */
nnd_status nnd_corr1d_build_nhwc_f16(const void* fmap1, const void* fmap2, int B, int C, int H, int W1, int W2,
                                     int num_levels, float* const* level, const int* pitch, nnd_stream_t stream);

/* Grouped variant.  Replaces GroupCorrBlock1D.corr / GeometryAwareCostVolume.build_cost_volume,
 * raft_stereo/cost_volume.py:113-128 and igev_stereo/cost_volume.py:81-98: group g contracts
 * channels [g*group_size, (g+1)*group_size) and divides by `scale_div` (sqrt(C) resp. sqrt(G)).
 *   level[l]: rows ordered [b][g][h][w1], width W2 >> l. */
nnd_status nnd_groupcorr_build(const float* fmap1, const float* fmap2, int B, int C, int H, int W1, int W2,
                               int num_groups, int group_size, float scale_div, int num_levels,
                               float* const* level, const int* pitch, nnd_stream_t stream);

/* avg_pool1d(x, 2) of `rows` rows: dst[r][j] = (src[r][2j] + src[r][2j+1]) / 2, j < src_width/2.
 * Replaces F.avg_pool1d at raft_stereo/cost_volume.py:33 for levels beyond the fused ones. */
nnd_status nnd_avgpool_pairs(const float* src, int src_width, int src_pitch, float* dst, int dst_pitch,
                             int64_t rows, nnd_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Fused pyramid lookup: all levels, all 2r+1 taps, NCHW output, one launch.
 * Replaces CorrBlock1D.__call__ raft_stereo/cost_volume.py:36-53 + linear_sampler
 * raft_stereo/utils.py:4-27.
 *   level[l] rows [b][h][w1] (rows = B*H*W1), widths width[l] >= 2, pitches pitch[l].
 *   coords (B,1,H,W1) -> out (B, num_levels*(2r+1), H, W1), channel = l*(2r+1)+k.
 * Index contract: t = clamp(x/(w-1),0,1)*(w-1) with IEEE fp32 div and mul; i0=floor(t), i1=ceil(t).
 * ---------------------------------------------------------------------------------------------- */
nnd_status nnd_corr1d_lookup(const float* const* level, const int* width, const int* pitch,
                             const float* coords, int B, int H, int W1, int num_levels, int radius,
                             float* out, nnd_stream_t stream);

/* The same lookup fused with the motion encoder's first layer, `F.relu(convc1(corr))` with convc1 a 1x1
 * convolution (nndepth/blocks/update_block.py:51,58): the (B, L*9, H, W) lookup tensor stays on chip.
 *   weight (L*9, c_out) row-major (= convc1.weight[:, :, 0, 0] transposed), bias (c_out) or NULL, relu 0/1;
 *   out (B, c_out, H, W).  radius 4, 4 levels.  precision NND_PREC_FP32: fp32 FFMA; NND_PREC_TF32
 *   (c_out <= 256): mma.sync TF32 with both operands rounded to nearest, weights resident in registers --
 *   the precision class cuDNN gives this layer under allow_tf32.  out_layout (tensor-core path only for
 *   1 and 2): 0 = fp32 (B, c_out, H, W); 1 = fp32 written (B, H, W, c_out), the layout cuDNN's tensor-core
 *   convolutions consume directly; 2 = that layout rounded to IEEE fp16 (the consumer runs as an fp16 convolution). */
nnd_status nnd_corr1d_lookup_conv1x1(const float* const* level, const int* width, const int* pitch,
                                     const float* coords, int B, int H, int W1, int num_levels, int radius,
                                     const float* weight, const float* bias, int c_out, int relu, int precision,
                                     int out_layout, void* out, nnd_stream_t stream);

/* Skewed copy of a CorrBlock1D pyramid and the fused lookup on it (private layout of this library; the reference's
 * volume is all-pairs (B,H,W1,W2), raft_stereo/cost_volume.py:55-61).  For every epipolar row (b, h), level l is stored
 * as a (W2_l x skew_pitch) matrix S[j][w1] = V_l[(b,h,w1)][w2] with j = ((w1 >> l) - w2) mod W2_l: the windows of
 * neighbouring pixels with similar disparity then share cache lines (a warp's 32 windows are ~12 rows of 128 contiguous
 * bytes), which removes the 64-byte-granule read amplification of one 40-byte window per 624-byte volume row.
 * skewed[l] needs B*H*width[l]*skew_pitch floats, skew_pitch >= W1 and a multiple of 4.
 * nnd_corr1d_lookup_conv1x1_skewed == nnd_corr1d_lookup_conv1x1 (TF32 operands, 4 levels, radius 4, c_out 256,
 * channels-last output: out_layout 1 = fp32, 2 = fp16) reading the skewed copy; results are bit-identical. */
nnd_status nnd_corr1d_skew(const float* const* level, const int* width, const int* pitch, int B, int H, int W1,
                           int num_levels, float* const* skewed, int skew_pitch, nnd_stream_t stream);
/* nnd_corr1d_lookup_skewed == nnd_corr1d_lookup (RAFT-Stereo form: one plane, radius 4; CorrBlock1D.__call__,
 * raft_stereo/cost_volume.py:36-53) reading the skewed copy: out (B, L*9, H, W1) fp32, bit-identical. */
nnd_status nnd_corr1d_lookup_skewed(const float* const* skewed, const int* width, int skew_pitch, const float* coords, int B,
                                    int H, int W1, int num_levels, int radius, float* out, nnd_stream_t stream);
nnd_status nnd_corr1d_lookup_conv1x1_skewed(const float* const* skewed, const int* width, int skew_pitch,
                                            const float* coords, int B, int H, int W1, int num_levels, int radius,
                                            const float* weight, const float* bias, int c_out, int relu, int out_layout,
                                            void* out, nnd_stream_t stream);

/* Backward passes for training (the reference path is differentiable; trainers: raft_trainer.py:242-259).
 *   nnd_corr1d_lookup_backward: gradient of the lookup w.r.t. the pyramid.  grad_out (B, L*(2r+1), H, W1);
 *     d_level[l] (rows = B*H*W1, pitch[l]) must be zero-initialised: each row receives its <= 2r+3 window
 *     entries, deterministic (a row belongs to one pixel).  The coordinates carry no gradient: the reference
 *     detaches them before every lookup (raft_stereo/model.py:131).
 *   nnd_avgpool_pairs_backward: backward of avg_pool1d(., 2) (raft_stereo/cost_volume.py:33), accumulating:
 *     d_src[r][2j] += d_dst[r][j] / 2, d_src[r][2j+1] += d_dst[r][j] / 2. */
nnd_status nnd_corr1d_lookup_backward(const float* grad_out, const float* coords, const int* width, const int* pitch,
                                      int B, int H, int W1, int num_levels, int radius, float* const* d_level,
                                      nnd_stream_t stream);
nnd_status nnd_avgpool_pairs_backward(const float* d_dst, int dst_width, int dst_pitch, float* d_src, int src_pitch,
                                      int64_t rows, nnd_stream_t stream);

/* Backward of the volume builds for the trainers (raft_trainer.py:242-259): the gradient of pyramid level 0 (rows
 * (b, g, h, w1) of `pitch` floats; g = 0 only for CorrBlock1D) contracted with the other feature map,
 *   which 0: d_fmap1[b,c,h,i] = 1/scale_div * sum_j dV[(b,g,h,i), j] * fmap2[b,c,h,j]      (f_other = fmap2)
 *   which 1: d_fmap2[b,c,h,j] = 1/scale_div * sum_i dV[(b,g,h,i), j] * fmap1[b,c,h,i]      (f_other = fmap1)
 * for c in group g = c / group_size -- the transpose of CorrBlock1D.corr / build_cost_volume
 * (raft_stereo/cost_volume.py:55-61, :113-128; igev_stereo/cost_volume.py:81-98).  Channels beyond
 * num_groups * group_size are not written (zero-fill d_fmap when they exist). */
nnd_status nnd_volume_grad(const float* d_level0, int pitch, const float* f_other, int B, int C, int H, int W1, int W2,
                           int num_groups, int group_size, float scale_div, int which, float* d_fmap, nnd_stream_t stream);

/* Debug/parity twin of the lookup: writes the int32 window indices instead of values.
 *   idx0, idx1: (num_levels, B*H*W1, 2r+1) int32. */
nnd_status nnd_corr1d_lookup_indices(const int* width, const float* coords, int B, int H, int W1,
                                     int num_levels, int radius, int32_t* idx0, int32_t* idx1,
                                     nnd_stream_t stream);

/* Grouped lookups over rows ordered [b][g][h][w1] (G groups share the pixel's coordinate).
 *   mode 0: IGEV dual lookup, GeometryAwareCostVolume.forward igev_stereo/cost_volume.py:54-79:
 *           two pyramids (feat, geo) -> out (B, L*2*G*T, H, W), channel l*(2GT) + src*(GT) + g*T + k.
 *           `level_b` = geometry pyramid.
 *   mode 1: GroupCorrBlock1D.__call__ raft_stereo/cost_volume.py:92-111: one pyramid, output is the
 *           reference's memory reinterpretation of [b][g][h][w][k] as (B,H,W,G*T) per level
 *           (cost_volume.py:108), then NCHW.  `level_b` ignored (may be NULL). */
nnd_status nnd_group_lookup(const float* const* level_a, const float* const* level_b, const int* width,
                            const int* pitch, const float* coords, int B, int G, int H, int W1,
                            int num_levels, int radius, int mode, float* out, nnd_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * IGEV geometry volume: regulariser output (B,G,D,H,W1) -> pyramid rows [b][g][h][w1] x D (+pooled).
 * Replaces the permute/reshape/avg_pool1d chain at igev_stereo/cost_volume.py:44-52.
 * ---------------------------------------------------------------------------------------------- */
nnd_status nnd_geo_transpose_pool(const float* geo, int B, int G, int D, int H, int W1, int num_levels,
                                  float* const* level, const int* pitch, nnd_stream_t stream);

/* Interleaved ("group innermost") IGEV pyramids -- the fast path for G == 8, D % 8 == 0, levels <= 4.
 * level[l] holds [b][h][w1][d_l][g] (one pixel = (D >> l) * 8 contiguous floats), so the eight radius-4
 * windows of a pixel are one contiguous run instead of eight scattered 40-byte pieces.
 *   nnd_gev_interleave_pool: one pass from either source layout to levels 0..num_levels-1, pooled exactly
 *     like avg_pool1d(., 2).  layout 0: reference-layout volume, rows [b][g][h][w1] of D floats with row
 *     pitch src_pitch (the output of nnd_groupcorr_build; igev_stereo/cost_volume.py:43-51).  layout 1:
 *     regulariser output (B, 8, D, H, W1) contiguous (igev_stereo/cost_volume.py:44-52); W1 % 4 == 0.
 *   nnd_gev_lookup: GeometryAwareCostVolume.forward igev_stereo/cost_volume.py:54-79 on those pyramids;
 *     coords (B,1,H,W1) -> out (B, L*2*8*9, H, W1), channel l*144 + src*72 + g*9 + k, radius 4. */
nnd_status nnd_gev_interleave_pool(const float* vol, int layout, int src_pitch, int B, int G, int D, int H,
                                   int W1, int num_levels, float* const* level, nnd_stream_t stream);
nnd_status nnd_gev_lookup(const float* const* level_feat, const float* const* level_geo, const float* coords,
                          int B, int G, int D, int H, int W1, int num_levels, int radius, float* out,
                          nnd_stream_t stream);

/* Soft-argmin: out[b,0,h,w] = -sum_d d * softmax_d(z[b,d,h,w]).  Single pass, online softmax.
 * Replaces F.softmax(dim=1) + regress_disparity, igev_stereo/model.py:145 and :92-95.
 *   z (B,D,H,W) -> out (B,1,H,W). */
nnd_status nnd_soft_argmin(const float* z, int B, int D, int H, int W, float* out, nnd_stream_t stream);

/* cv_squeezer + soft-argmin in one pass over the interleaved level-0 geometry volume:
 *   cost[b,d,h,w] = bias + sum_{g,kd,kh,kw} weight[0,g,kd,kh,kw] * geo[b,g,d+kd-1,h+kh-1,w+kw-1]  (zero padded)
 *   out[b,0,h,w]  = -sum_d d * softmax_d(cost[b,:,h,w])
 * Replaces nn.Conv3d(8, 1, 3, 1, 1) on geo_aware_cv[0].reshape(...).permute(0,1,4,2,3), squeeze, F.softmax and
 * regress_disparity: igev_stereo/model.py:65, 143-146, 92-95.
 *   geo_level0: level 0 written by nnd_gev_interleave_pool ([b][h][w1][d][g], G = 8); weight: the Conv3d
 *   weight (1,8,3,3,3) contiguous, device memory; bias: 1 float in device memory or NULL; D <= 512;
 *   out (B,1,H,W1); cost_out: optional (B,D,H,W1) copy of the squeezed cost (parity tests), or NULL. */
nnd_status nnd_gev_squeeze_soft_argmin(const float* geo_level0, const float* weight, const float* bias, int B,
                                       int G, int D, int H, int W1, float* out, float* cost_out,
                                       nnd_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * CREStereo AGCL.  fmap1/fmap2 (N,C,H,W) NCHW, C % 4 == 0; flow (N,2,H,W) (x, y);
 * out (N, 36, H, W), channel g*9 + k; small_patch: 0 = 1x9 window, 1 = 3x3 window.
 *   nnd_agcl_offset: AGCL.corr_att_offset nndepth/models/cre_stereo/cost_volume.py:81-154 (the
 *     optional attention is applied by the caller beforehand); extra_offset (N,18,H,W).
 *   nnd_agcl_iter:   AGCL.corr_iter :54-79 + get_correlation :28-52 (warp by flow, replicate-padded
 *     local window).
 * Samplers follow bilinear_sampler / bilinear_grid_sample cre_stereo/utils.py:5-107 (zero padding,
 * align_corners=True, fp32 normalise/denormalise round trip kept for bit-exact corner indices).
 * ---------------------------------------------------------------------------------------------- */
nnd_status nnd_agcl_offset(const float* fmap1, const float* fmap2, const float* flow,
                           const float* extra_offset, int N, int C, int H, int W, int small_patch,
                           float* out, nnd_stream_t stream);
nnd_status nnd_agcl_iter(const float* fmap1, const float* fmap2, const float* flow, int N, int C, int H,
                         int W, int small_patch, float* out, nnd_stream_t stream);

/* Channels-last fast path of AGCL (C % 16 == 0, C <= 512).  The maps of an AGCL object are fixed while
 * it is called 6-12 times per cascade scale (cre_stereo/model.py:198-275), so the caller stages them
 * once as (N,H,W,C) with nnd_nchw_to_nhwc and every call gathers whole channel vectors (one bilinear
 * corner = C contiguous floats).  Same semantics, outputs and reference citations as the NCHW entry
 * points above.  nnd_agcl_iter_nhwc needs a caller-provided workspace of N*H*W*C floats for the
 * flow-warped right map (the tensor the reference materialises at cost_volume.py:57-59). */
nnd_status nnd_nchw_to_nhwc(const float* src, int N, int C, int H, int W, float* dst, nnd_stream_t stream);
nnd_status nnd_agcl_offset_nhwc(const float* fmap1_nhwc, const float* fmap2_nhwc, const float* flow,
                                const float* extra_offset, int N, int C, int H, int W, int small_patch,
                                float* out, nnd_stream_t stream);
nnd_status nnd_agcl_iter_nhwc(const float* fmap1_nhwc, const float* fmap2_nhwc, const float* flow, int N,
                              int C, int H, int W, int small_patch, float* warped_ws, float* out,
                              nnd_stream_t stream);

/* Backward of the channels-last AGCL entry points for the reference's trainers (the forward they differentiate:
 * cre_stereo/cost_volume.py:54-154; bilinear sampling utils.py:34-107).  grad_out is (N,36,H,W).  d_fmap1 (N,H,W,C) is
 * written; d_fmap2 (N,H,W,C) -- and d_warped_ws for iter mode -- are ACCUMULATED with atomics: the caller zero-fills
 * them.  d_flow (N,2,H,W) and d_extra (N,18,H,W) may be NULL.  nnd_agcl_warp_nhwc materialises the flow-warped right
 * map (cost_volume.py:57-59) that iter mode's backward re-reads.  nnd_agcl_iter_backward_nhwc with d_fmap2 == NULL
 * computes d_fmap1 only: the reference detaches the warped map (manual_pad, cre_stereo/utils.py:29-31), so its iter
 * mode trains the left features alone. */
nnd_status nnd_agcl_warp_nhwc(const float* fmap2_nhwc, const float* flow, int N, int C, int H, int W, float* warped,
                              nnd_stream_t stream);
nnd_status nnd_agcl_offset_backward_nhwc(const float* fmap1_nhwc, const float* fmap2_nhwc, const float* flow,
                                         const float* extra_offset, const float* grad_out, int N, int C, int H, int W,
                                         int small_patch, float* d_fmap1, float* d_fmap2, float* d_flow, float* d_extra,
                                         nnd_stream_t stream);
nnd_status nnd_agcl_iter_backward_nhwc(const float* fmap1_nhwc, const float* fmap2_nhwc, const float* flow,
                                       const float* warped, const float* grad_out, int N, int C, int H, int W,
                                       int small_patch, float* d_fmap1, float* d_fmap2, float* d_flow, float* d_warped_ws,
                                       nnd_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Convex upsampling of the coarse disparity by the learned 9-neighbour mask, one pass.
 * Replaces RAFTStereo.convex_upsample nndepth/models/raft_stereo/model.py:93-105 (identical code in
 * cre_stereo/model.py:110-122 and igev_stereo/model.py:103-115): softmax over the 9 neighbours,
 * F.unfold(rate * flow, 3x3, padding 1), multiply, sum, pixel shuffle.
 *   flow (N,1,H,W), mask (N, 9*rate*rate, H, W) -> out (N, 1, rate*H, rate*W); rate in {2, 4, 8}.
 *   mask_scale multiplies the mask logits first: pass 0.25 to fold the update block's `0.25 * mask`
 *   (blocks/update_block.py:110) into this pass, 1.0 for an already scaled mask.
 *   mask_bias (9*rate*rate) or NULL is added to the logits before the scale: the bias of the mask head's last
 *   1x1 convolution, so that convolution can run bias-free (one pass over the mask less).
 *   mask_layout: 0 = fp32 (N, 9*rate*rate, H, W); 1 = fp32 stored (N, H, W, 9*rate*rate) (rate 8 only) -- the
 *   memory format cuDNN returns when the hidden state feeding the mask head is channels-last; 2 = the same
 *   channels-last layout in IEEE fp16 (mask head run as fp16 convolutions).
 * ---------------------------------------------------------------------------------------------- */
nnd_status nnd_convex_upsample(const float* flow, const void* mask, const float* mask_bias, int N, int H, int W,
                               int rate, float mask_scale, int mask_layout, float* out, nnd_stream_t stream);

/* Single-flow-channel convolutions of the update block (stereo: flow_channel = 1), fp32 FFMA.
 *   nnd_flow_conv7x7_relu: relu(convf1(flow)), BasicMotionEncoder blocks/update_block.py:53,60.
 *     flow (N,1,H,W); weight_t (49, c_out) = the (c_out,1,7,7) filter bank transposed (tap-major); bias (c_out);
 *     out channels-last (N,H,W,c_out), fp32 or (out_f16 != 0) IEEE fp16; padding 3.
 *   nnd_flow_head_tail: FlowHead.conv2 blocks/update_block.py:23,36 on a channels-last x (N,H,W,C) (fp32, or IEEE fp16
 *     when x_f16), C in
 *     {128, 256, 512}; weight (1,C,3,3); bias 1 float or NULL; delta (N,1,H,W) or NULL.  With coords_in the
 *     refinement-loop update raft_stereo/model.py:132-134 is fused: coords_out = coords_in + delta and, if
 *     flow_out, flow_out = coords_out - org (all (N,1,H,W); coords_out may alias coords_in). */
nnd_status nnd_flow_conv7x7_relu(const float* flow, const float* weight_t, const float* bias, int N, int H, int W,
                                 int c_out, void* out, int out_f16, nnd_stream_t stream);
/* torch.cat([a, b], dim=1) of two channels-last maps (blocks/update_block.py:62) converted to IEEE fp16 on the way:
 * a (pixels, c_a), b (pixels, c_b), each fp32 or (x_f16 != 0) fp16 -> out (pixels, c_a + c_b) fp16; c_a, c_b % 4 == 0. */
nnd_status nnd_nhwc_cat_f16(const void* a, int a_f16, int c_a, const void* b, int b_f16, int c_b, long long pixels,
                            void* out, nnd_stream_t stream);
nnd_status nnd_flow_head_tail(const void* x, int x_f16, const float* weight, const float* bias, int N, int C, int H,
                              int W, float* delta, const float* coords_in, const float* org, float* coords_out,
                              float* flow_out, nnd_stream_t stream);

/* Fused, channels-last glue of the separable ConvGRU (nndepth/blocks/gru.py:5-37) around its weight-split
 * TF32 convolutions: conv([RN_tf32(x) ; RN_tf32(x)], [w_hi ; w_lo]) -- fp32-exact weights on the tensor cores;
 * plain TF32 weights leave the 0.01 px parity bar.  One staging buffer S (N, H*W, ctot), rows
 * [ RN(h) (ch) | RN(x) (cx) | RN(h) | RN(x) ], ctot = 2*(ch + cx), is the NHWC input of all four convolutions of
 * an iteration.
 *   nnd_gru_stage:  src (N, C, H*W) NCHW, or (N, H*W, C) when src_channels_last != 0 -> RN_tf32(src) at channel
 *                   offsets `off` and `off + ctot/2` of S.
 *   nnd_gru_gate_r: zr_pre (pixels, 2ch) NHWC conv output [z | r], bias_zr (2ch), h (pixels, ch)
 *                   -> z = sigmoid(z_pre + b) (pixels, ch);  S.h <- RN(sigmoid(r_pre + b) * h) (both copies).
 *   nnd_gru_gate_h: q_pre (pixels, ch), bias_q (ch), z, h -> h <- (1 - z) * h + z * tanh(q_pre + b) in place;
 *                   S.h <- RN(h) (both copies). */
nnd_status nnd_gru_stage(const float* src, int src_channels_last, int N, int C, long long hw, float* S, int ctot,
                         int off, nnd_stream_t stream);
nnd_status nnd_gru_gate_r(const float* zr_pre, const float* bias_zr, const float* h, long long pixels, int ch,
                          float* z, float* S, int ctot, nnd_stream_t stream);
nnd_status nnd_gru_gate_h(const float* q_pre, const float* bias_q, const float* z, long long pixels, int ch,
                          float* h, float* S, int ctot, nnd_stream_t stream);

/* The same glue for the fp16 form of the recurrence (kind::f16 tensor-core products at twice the TF32 rate, same
 * 10-bit operand mantissa): staging buffer S16, split weights and the convolutions' pre-activations are IEEE
 * fp16 (passed as void*), h / z / biases stay fp32; conversions round to nearest and saturate.
 *   nnd_gru_stage_f16: src_kind 0 = fp32 (N,C,HW), 1 = fp32 channels-last, 2 = fp16 channels-last.
 *   nnd_gru_gate_h_f16: h16, when not NULL, also receives the new hidden state as dense channels-last fp16. */
nnd_status nnd_gru_stage_f16(const void* src, int src_kind, int N, int C, long long hw, void* S16, int ctot, int off,
                             nnd_stream_t stream);
nnd_status nnd_gru_gate_r_f16(const void* zr_pre16, const float* bias_zr, const float* h, long long pixels, int ch,
                              float* z, void* S16, int ctot, nnd_stream_t stream);
nnd_status nnd_gru_gate_h_f16(const void* q_pre16, const float* bias_q, const float* z, long long pixels, int ch,
                              float* h, void* S16, int ctot, void* h16, nnd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* NNDEPTH_B200_H */

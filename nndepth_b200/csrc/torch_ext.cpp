// Thin PyTorch C++ extension over the C ABI of include/nndepth_b200.h:
// TORCH_LIBRARY(nndepth_b200, ...) operators that check device / dtype / contiguity / shapes, take the CURRENT CUDA stream
// of the tensors' device, allocate outputs through the caching allocator and hand raw pointers to libnndepth_b200.so.
// No arithmetic lives here.  A non-zero status becomes a RuntimeError carrying nnd_last_error_string().
//
// Pyramids cross the boundary as ONE flat fp32 buffer (corr.py: PyramidStorage -- level l is a (rows, pitch_l) matrix,
// pitch_l = roundup4(width0 >> l), levels concatenated; igev.py: InterleavedPyramid -- level l holds
// pixels * (D >> l) * 8 floats); the level tables the C ABI wants are rebuilt here from (rows, width0, num_levels).
#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include <array>
#include <tuple>
#include <vector>

#include "nndepth_b200.h"

namespace {

using at::Tensor;
using c10::optional;

inline nnd_stream_t current_stream(const Tensor& t) {
  return reinterpret_cast<nnd_stream_t>(c10::cuda::getCurrentCUDAStream(t.get_device()).stream());
}

inline void check_status(nnd_status st, const char* what) {
  TORCH_CHECK(st == NND_OK, what, " failed (status ", st, "): ", nnd_last_error_string());
}

inline void check_cuda(const Tensor& t, const char* name, at::ScalarType dtype = at::kFloat) {
  TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor: nndepth_b200 has no CPU path (got device ", t.device(), ")");
  TORCH_CHECK(t.scalar_type() == dtype, name, " must be ", dtype, ", got ", t.scalar_type());
  TORCH_CHECK(t.is_contiguous(), name, " must be contiguous (shape ", t.sizes(), ", strides ", t.strides(), ")");
}

inline void same_device(const Tensor& a, const Tensor& b, const char* what) {
  TORCH_CHECK(a.device() == b.device(), what, ": tensors live on different devices (", a.device(), " vs ", b.device(), ")");
}

struct LevelTable {
  std::array<float*, NND_MAX_LEVELS> ptr{};
  std::array<int, NND_MAX_LEVELS> width{};
  std::array<int, NND_MAX_LEVELS> pitch{};
  int64_t floats = 0;
};

// row-pitched pyramid (PyramidStorage)
LevelTable row_levels(const Tensor& buffer, int64_t rows, int64_t width0, int64_t num_levels, const char* who) {
  TORCH_CHECK(num_levels >= 1 && num_levels <= NND_MAX_LEVELS, who, ": num_levels must be in [1, ", NND_MAX_LEVELS, "], got ",
              num_levels);
  LevelTable t;
  float* base = buffer.data_ptr<float>();
  for (int l = 0; l < num_levels; ++l) {
    t.width[l] = static_cast<int>(width0 >> l);
    TORCH_CHECK(t.width[l] >= 1, who, ": a ", num_levels, "-level pyramid of width ", width0, " has an empty level");
    t.pitch[l] = nnd_row_pitch(t.width[l]);
    t.ptr[l] = base + t.floats;
    t.floats += rows * t.pitch[l];
  }
  TORCH_CHECK(buffer.numel() >= t.floats, who, ": pyramid buffer holds ", buffer.numel(), " floats, needs ", t.floats);
  return t;
}

// pixel-major, group-innermost pyramid (InterleavedPyramid, G = 8)
LevelTable interleaved_levels(const Tensor& buffer, int64_t pixels, int64_t depth, int64_t num_levels, const char* who) {
  TORCH_CHECK(num_levels >= 1 && num_levels <= 4, who, ": num_levels must be in [1, 4], got ", num_levels);
  LevelTable t;
  float* base = buffer.data_ptr<float>();
  for (int l = 0; l < num_levels; ++l) {
    t.width[l] = static_cast<int>(depth >> l);
    t.ptr[l] = base + t.floats;
    t.floats += pixels * t.width[l] * 8;
  }
  TORCH_CHECK(buffer.numel() >= t.floats, who, ": interleaved buffer holds ", buffer.numel(), " floats, needs ", t.floats);
  return t;
}

inline std::array<int64_t, 4> nchw(const Tensor& t, const char* name) {
  TORCH_CHECK(t.dim() == 4, name, " must be 4-D, got ", t.sizes());
  return {t.size(0), t.size(1), t.size(2), t.size(3)};
}

// ---------------------------------------------------------------------------------------------------------------
// RAFT-Stereo: CorrBlock1D / GroupCorrBlock1D  (raft_stereo/cost_volume.py:7-128)
// ---------------------------------------------------------------------------------------------------------------
void corr1d_build(const Tensor& f1, const Tensor& f2, Tensor pyramid, int64_t num_levels, int64_t precision) {
  check_cuda(f1, "fmap1"); check_cuda(f2, "fmap2"); check_cuda(pyramid, "pyramid");
  same_device(f1, f2, "corr1d_build"); same_device(f1, pyramid, "corr1d_build");
  auto s1 = nchw(f1, "fmap1"), s2 = nchw(f2, "fmap2");
  TORCH_CHECK(s1[0] == s2[0] && s1[1] == s2[1] && s1[2] == s2[2], "fmap1 ", f1.sizes(), " and fmap2 ", f2.sizes(),
              " must agree in batch, channels and height");
  c10::cuda::CUDAGuard guard(f1.device());
  LevelTable t = row_levels(pyramid, s1[0] * s1[2] * s1[3], s2[3], num_levels, "corr1d_build");
  check_status(nnd_corr1d_build(f1.data_ptr<float>(), f2.data_ptr<float>(), s1[0], s1[1], s1[2], s1[3], s2[3], num_levels,
                                precision, t.ptr.data(), t.pitch.data(), current_stream(f1)),
               "nnd_corr1d_build");
}

// the same build from fp16 channels-last feature maps (N, C, H, W logical; N, H, W, C in memory), read where they lie
void corr1d_build_nhwc_f16(const Tensor& f1, const Tensor& f2, Tensor pyramid, int64_t num_levels) {
  check_cuda(pyramid, "pyramid");
  for (const Tensor* f : {&f1, &f2}) {
    TORCH_CHECK(f->is_cuda() && f->scalar_type() == at::kHalf && f->dim() == 4 &&
                    f->is_contiguous(at::MemoryFormat::ChannelsLast),
                "corr1d_build_nhwc_f16: feature maps must be dense fp16 channels-last CUDA tensors");
  }
  same_device(f1, f2, "corr1d_build_nhwc_f16"); same_device(f1, pyramid, "corr1d_build_nhwc_f16");
  auto s1 = nchw(f1, "fmap1"), s2 = nchw(f2, "fmap2");
  TORCH_CHECK(s1[0] == s2[0] && s1[1] == s2[1] && s1[2] == s2[2], "fmap1 ", f1.sizes(), " and fmap2 ", f2.sizes(),
              " must agree in batch, channels and height");
  c10::cuda::CUDAGuard guard(f1.device());
  LevelTable t = row_levels(pyramid, s1[0] * s1[2] * s1[3], s2[3], num_levels, "corr1d_build_nhwc_f16");
  check_status(nnd_corr1d_build_nhwc_f16(f1.data_ptr(), f2.data_ptr(), s1[0], s1[1], s1[2], s1[3], s2[3], num_levels,
                                         t.ptr.data(), t.pitch.data(), current_stream(f1)),
               "nnd_corr1d_build_nhwc_f16");
}

void groupcorr_build(const Tensor& f1, const Tensor& f2, Tensor pyramid, int64_t num_groups, int64_t group_size,
                     double scale_div, int64_t num_levels) {
  check_cuda(f1, "fmap1"); check_cuda(f2, "fmap2"); check_cuda(pyramid, "pyramid");
  same_device(f1, f2, "groupcorr_build"); same_device(f1, pyramid, "groupcorr_build");
  auto s1 = nchw(f1, "fmap1"), s2 = nchw(f2, "fmap2");
  TORCH_CHECK(s1[0] == s2[0] && s1[1] == s2[1] && s1[2] == s2[2], "fmap1 and fmap2 must agree in batch, channels and height");
  c10::cuda::CUDAGuard guard(f1.device());
  LevelTable t = row_levels(pyramid, s1[0] * num_groups * s1[2] * s1[3], s2[3], num_levels, "groupcorr_build");
  check_status(nnd_groupcorr_build(f1.data_ptr<float>(), f2.data_ptr<float>(), s1[0], s1[1], s1[2], s1[3], s2[3], num_groups,
                                   group_size, static_cast<float>(scale_div), num_levels, t.ptr.data(), t.pitch.data(),
                                   current_stream(f1)),
               "nnd_groupcorr_build");
}

Tensor avgpool_pairs(const Tensor& src, int64_t src_width) {
  check_cuda(src, "src");
  TORCH_CHECK(src.dim() == 2, "src must be (rows, pitch)");
  c10::cuda::CUDAGuard guard(src.device());
  const int pitch = nnd_row_pitch(static_cast<int>(src_width / 2));
  Tensor dst = at::empty({src.size(0), pitch}, src.options());
  check_status(nnd_avgpool_pairs(src.data_ptr<float>(), src_width, src.size(1), dst.data_ptr<float>(), pitch, src.size(0),
                                 current_stream(src)),
               "nnd_avgpool_pairs");
  return dst;
}

Tensor corr1d_lookup(const Tensor& pyramid, int64_t width0, const Tensor& coords, int64_t num_levels, int64_t radius) {
  check_cuda(pyramid, "pyramid"); check_cuda(coords, "coords");
  same_device(pyramid, coords, "corr1d_lookup");
  auto s = nchw(coords, "coords");
  TORCH_CHECK(s[1] == 1, "coords must be (B, 1, H, W), got ", coords.sizes());
  c10::cuda::CUDAGuard guard(coords.device());
  LevelTable t = row_levels(pyramid, s[0] * s[2] * s[3], width0, num_levels, "corr1d_lookup");
  Tensor out = at::empty({s[0], num_levels * (2 * radius + 1), s[2], s[3]}, coords.options());
  check_status(nnd_corr1d_lookup(t.ptr.data(), t.width.data(), t.pitch.data(), coords.data_ptr<float>(), s[0], s[2], s[3],
                                 num_levels, radius, out.data_ptr<float>(), current_stream(coords)),
               "nnd_corr1d_lookup");
  return out;
}

Tensor corr1d_lookup_conv1x1(const Tensor& pyramid, int64_t width0, const Tensor& coords, int64_t num_levels, int64_t radius,
                             const Tensor& weight_t, const optional<Tensor>& bias, bool relu, int64_t precision,
                             int64_t out_layout) {
  check_cuda(pyramid, "pyramid"); check_cuda(coords, "coords"); check_cuda(weight_t, "weight_t");
  same_device(pyramid, coords, "corr1d_lookup_conv1x1"); same_device(weight_t, coords, "corr1d_lookup_conv1x1");
  auto s = nchw(coords, "coords");
  TORCH_CHECK(s[1] == 1, "coords must be (B, 1, H, W), got ", coords.sizes());
  TORCH_CHECK(weight_t.dim() == 2 && weight_t.size(0) == num_levels * (2 * radius + 1), "weight_t must be (",
              num_levels * (2 * radius + 1), ", c_out), got ", weight_t.sizes());
  const float* bias_ptr = nullptr;
  if (bias.has_value()) {
    check_cuda(*bias, "bias");
    TORCH_CHECK(bias->numel() == weight_t.size(1), "bias must have c_out = ", weight_t.size(1), " elements");
    bias_ptr = bias->data_ptr<float>();
  }
  TORCH_CHECK(out_layout >= 0 && out_layout <= 2, "out_layout must be 0 (NCHW fp32), 1 (NHWC fp32) or 2 (NHWC fp16)");
  c10::cuda::CUDAGuard guard(coords.device());
  LevelTable t = row_levels(pyramid, s[0] * s[2] * s[3], width0, num_levels, "corr1d_lookup_conv1x1");
  const int64_t c_out = weight_t.size(1);
  Tensor out;
  if (out_layout == 0) {
    out = at::empty({s[0], c_out, s[2], s[3]}, coords.options());
  } else {
    // (B, H, W, c_out) in memory, returned with NCHW shape and channels-last strides
    out = at::empty({s[0], s[2], s[3], c_out}, coords.options().dtype(out_layout == 2 ? at::kHalf : at::kFloat)).permute({0, 3, 1, 2});
  }
  check_status(nnd_corr1d_lookup_conv1x1(t.ptr.data(), t.width.data(), t.pitch.data(), coords.data_ptr<float>(), s[0], s[2],
                                         s[3], num_levels, radius, weight_t.data_ptr<float>(), bias_ptr, c_out, relu ? 1 : 0,
                                         precision, out_layout, out.data_ptr(), current_stream(coords)),
               "nnd_corr1d_lookup_conv1x1");
  return out;
}

// skewed copy of a row-layout pyramid: one flat buffer, level l = B*H*(W2 >> l) rows of roundup4(W1) floats
Tensor corr1d_skew(const Tensor& pyramid, int64_t B, int64_t H, int64_t W1, int64_t W2, int64_t num_levels) {
  check_cuda(pyramid, "pyramid");
  c10::cuda::CUDAGuard guard(pyramid.device());
  LevelTable t = row_levels(pyramid, B * H * W1, W2, num_levels, "corr1d_skew");
  const int P1 = nnd_row_pitch(static_cast<int>(W1));
  int64_t floats = 0;
  for (int l = 0; l < num_levels; ++l) floats += B * H * t.width[l] * P1;
  Tensor skew = at::empty({floats}, pyramid.options());
  std::array<float*, NND_MAX_LEVELS> dst{};
  int64_t off = 0;
  for (int l = 0; l < num_levels; ++l) {
    dst[l] = skew.data_ptr<float>() + off;
    off += B * H * t.width[l] * P1;
  }
  check_status(nnd_corr1d_skew(t.ptr.data(), t.width.data(), t.pitch.data(), B, H, W1, num_levels, dst.data(), P1,
                               current_stream(pyramid)),
               "nnd_corr1d_skew");
  return skew;
}

Tensor corr1d_lookup_conv1x1_skewed(const Tensor& skew, int64_t width0, const Tensor& coords, int64_t num_levels, int64_t radius,
                                    const Tensor& weight_t, const optional<Tensor>& bias, bool relu, int64_t out_layout) {
  check_cuda(skew, "skewed pyramid"); check_cuda(coords, "coords"); check_cuda(weight_t, "weight_t");
  same_device(skew, coords, "corr1d_lookup_conv1x1_skewed"); same_device(weight_t, coords, "corr1d_lookup_conv1x1_skewed");
  auto s = nchw(coords, "coords");
  TORCH_CHECK(s[1] == 1, "coords must be (B, 1, H, W), got ", coords.sizes());
  TORCH_CHECK(weight_t.dim() == 2 && weight_t.size(0) == num_levels * (2 * radius + 1) && weight_t.size(1) == 256,
              "weight_t must be (", num_levels * (2 * radius + 1), ", 256), got ", weight_t.sizes());
  TORCH_CHECK(out_layout == 1 || out_layout == 2, "out_layout must be 1 (NHWC fp32) or 2 (NHWC fp16)");
  const float* bias_ptr = nullptr;
  if (bias.has_value()) {
    check_cuda(*bias, "bias");
    TORCH_CHECK(bias->numel() == 256, "bias must have 256 elements");
    bias_ptr = bias->data_ptr<float>();
  }
  c10::cuda::CUDAGuard guard(coords.device());
  const int64_t B = s[0], H = s[2], W1 = s[3];
  const int P1 = nnd_row_pitch(static_cast<int>(W1));
  std::array<const float*, NND_MAX_LEVELS> lv{};
  std::array<int, NND_MAX_LEVELS> width{};
  int64_t off = 0;
  for (int l = 0; l < num_levels; ++l) {
    width[l] = static_cast<int>(width0 >> l);
    lv[l] = skew.data_ptr<float>() + off;
    off += B * H * width[l] * P1;
  }
  TORCH_CHECK(skew.numel() >= off, "skewed pyramid holds ", skew.numel(), " floats, needs ", off);
  Tensor out = at::empty({B, H, W1, 256}, coords.options().dtype(out_layout == 2 ? at::kHalf : at::kFloat)).permute({0, 3, 1, 2});
  check_status(nnd_corr1d_lookup_conv1x1_skewed(lv.data(), width.data(), P1, coords.data_ptr<float>(), B, H, W1, num_levels,
                                                radius, weight_t.data_ptr<float>(), bias_ptr, 256, relu ? 1 : 0, out_layout,
                                                out.data_ptr(), current_stream(coords)),
               "nnd_corr1d_lookup_conv1x1_skewed");
  return out;
}

Tensor corr1d_lookup_skewed(const Tensor& skew, int64_t width0, const Tensor& coords, int64_t num_levels, int64_t radius) {
  check_cuda(skew, "skewed pyramid"); check_cuda(coords, "coords");
  same_device(skew, coords, "corr1d_lookup_skewed");
  auto s = nchw(coords, "coords");
  TORCH_CHECK(s[1] == 1, "coords must be (B, 1, H, W), got ", coords.sizes());
  TORCH_CHECK(num_levels >= 1 && num_levels <= NND_MAX_LEVELS, "num_levels out of range");
  c10::cuda::CUDAGuard guard(coords.device());
  const int64_t B = s[0], H = s[2], W1 = s[3];
  const int P1 = nnd_row_pitch(static_cast<int>(W1));
  std::array<const float*, NND_MAX_LEVELS> lv{};
  std::array<int, NND_MAX_LEVELS> width{};
  int64_t off = 0;
  for (int l = 0; l < num_levels; ++l) {
    width[l] = static_cast<int>(width0 >> l);
    lv[l] = skew.data_ptr<float>() + off;
    off += B * H * width[l] * P1;
  }
  TORCH_CHECK(skew.numel() >= off, "skewed pyramid holds ", skew.numel(), " floats, needs ", off);
  Tensor out = at::empty({B, num_levels * (2 * radius + 1), H, W1}, coords.options());
  check_status(nnd_corr1d_lookup_skewed(lv.data(), width.data(), P1, coords.data_ptr<float>(), B, H, W1, num_levels, radius,
                                        out.data_ptr<float>(), current_stream(coords)),
               "nnd_corr1d_lookup_skewed");
  return out;
}

Tensor corr1d_lookup_backward(const Tensor& grad_out, const Tensor& coords, int64_t width0, int64_t num_levels, int64_t radius) {
  check_cuda(grad_out, "grad_out"); check_cuda(coords, "coords");
  same_device(grad_out, coords, "corr1d_lookup_backward");
  auto s = nchw(coords, "coords");
  TORCH_CHECK(grad_out.dim() == 4 && grad_out.size(0) == s[0] && grad_out.size(1) == num_levels * (2 * radius + 1) &&
                  grad_out.size(2) == s[2] && grad_out.size(3) == s[3],
              "grad_out ", grad_out.sizes(), " does not match the lookup output of coords ", coords.sizes());
  c10::cuda::CUDAGuard guard(coords.device());
  const int64_t rows = s[0] * s[2] * s[3];
  int64_t floats = 0;
  for (int l = 0; l < num_levels; ++l) floats += rows * nnd_row_pitch(static_cast<int>(width0 >> l));
  Tensor d = at::zeros({floats}, coords.options());
  LevelTable t = row_levels(d, rows, width0, num_levels, "corr1d_lookup_backward");
  check_status(nnd_corr1d_lookup_backward(grad_out.data_ptr<float>(), coords.data_ptr<float>(), t.width.data(), t.pitch.data(),
                                          s[0], s[2], s[3], num_levels, radius, t.ptr.data(), current_stream(coords)),
               "nnd_corr1d_lookup_backward");
  return d;
}

// un-pool every level of a pyramid-shaped gradient buffer in place, coarsest level first (avg_pool1d backward)
void pyramid_unpool_(Tensor d_pyramid, int64_t rows, int64_t width0, int64_t num_levels) {
  check_cuda(d_pyramid, "d_pyramid");
  c10::cuda::CUDAGuard guard(d_pyramid.device());
  LevelTable t = row_levels(d_pyramid, rows, width0, num_levels, "pyramid_unpool_");
  for (int l = static_cast<int>(num_levels) - 1; l > 0; --l)
    check_status(nnd_avgpool_pairs_backward(t.ptr[l], t.width[l], t.pitch[l], t.ptr[l - 1], t.pitch[l - 1], rows,
                                            current_stream(d_pyramid)),
                 "nnd_avgpool_pairs_backward");
}

// d_fmap of one side of the build (which 0: fmap1, 1: fmap2) from the un-pooled pyramid-gradient buffer (its level 0)
Tensor volume_grad(const Tensor& d_pyramid, const Tensor& f_other, int64_t W1, int64_t W2, int64_t num_groups, int64_t group_size,
                   double scale_div, int64_t which) {
  check_cuda(d_pyramid, "d_pyramid"); check_cuda(f_other, "f_other");
  same_device(d_pyramid, f_other, "volume_grad");
  auto s = nchw(f_other, "f_other");
  const int64_t B = s[0], C = s[1], H = s[2];
  TORCH_CHECK(s[3] == (which == 0 ? W2 : W1), "volume_grad: f_other has width ", s[3], ", expected ", (which == 0 ? W2 : W1));
  const int pitch = nnd_row_pitch(static_cast<int>(W2));
  TORCH_CHECK(d_pyramid.numel() >= B * num_groups * H * W1 * pitch, "volume_grad: gradient buffer smaller than level 0");
  c10::cuda::CUDAGuard guard(f_other.device());
  const bool partial = num_groups * group_size < C;
  Tensor out = partial ? at::zeros({B, C, H, which == 0 ? W1 : W2}, f_other.options())
                       : at::empty({B, C, H, which == 0 ? W1 : W2}, f_other.options());
  check_status(nnd_volume_grad(d_pyramid.data_ptr<float>(), pitch, f_other.data_ptr<float>(), B, C, H, W1, W2, num_groups,
                               group_size, static_cast<float>(scale_div), which, out.data_ptr<float>(), current_stream(f_other)),
               "nnd_volume_grad");
  return out;
}

std::tuple<Tensor, Tensor> corr1d_lookup_indices(at::IntArrayRef widths, const Tensor& coords, int64_t num_levels,
                                                 int64_t radius) {
  check_cuda(coords, "coords");
  auto s = nchw(coords, "coords");
  TORCH_CHECK(static_cast<int64_t>(widths.size()) >= num_levels, "widths must list num_levels entries");
  std::vector<int> w(widths.begin(), widths.begin() + num_levels);
  c10::cuda::CUDAGuard guard(coords.device());
  auto opt = coords.options().dtype(at::kInt);
  Tensor i0 = at::empty({num_levels, s[0] * s[2] * s[3], 2 * radius + 1}, opt), i1 = at::empty_like(i0);
  check_status(nnd_corr1d_lookup_indices(w.data(), coords.data_ptr<float>(), s[0], s[2], s[3], num_levels, radius,
                                         i0.data_ptr<int32_t>(), i1.data_ptr<int32_t>(), current_stream(coords)),
               "nnd_corr1d_lookup_indices");
  return {i0, i1};
}

Tensor group_lookup(const Tensor& pyr_a, const optional<Tensor>& pyr_b, int64_t width0, const Tensor& coords, int64_t G,
                    int64_t num_levels, int64_t radius, int64_t mode) {
  check_cuda(pyr_a, "pyramid"); check_cuda(coords, "coords");
  same_device(pyr_a, coords, "group_lookup");
  auto s = nchw(coords, "coords");
  TORCH_CHECK(s[1] == 1, "coords must be (B, 1, H, W), got ", coords.sizes());
  c10::cuda::CUDAGuard guard(coords.device());
  const int64_t rows = s[0] * G * s[2] * s[3];
  LevelTable ta = row_levels(pyr_a, rows, width0, num_levels, "group_lookup"), tb;
  const int64_t n_src = pyr_b.has_value() ? 2 : 1;
  if (pyr_b.has_value()) {
    check_cuda(*pyr_b, "second pyramid");
    tb = row_levels(*pyr_b, rows, width0, num_levels, "group_lookup");
  }
  Tensor out = at::empty({s[0], num_levels * n_src * G * (2 * radius + 1), s[2], s[3]}, coords.options());
  check_status(nnd_group_lookup(ta.ptr.data(), pyr_b.has_value() ? tb.ptr.data() : nullptr, ta.width.data(), ta.pitch.data(),
                                coords.data_ptr<float>(), s[0], G, s[2], s[3], num_levels, radius, mode, out.data_ptr<float>(),
                                current_stream(coords)),
               "nnd_group_lookup");
  return out;
}

// ---------------------------------------------------------------------------------------------------------------
// IGEV-Stereo: GeometryAwareCostVolume + soft-argmin  (igev_stereo/cost_volume.py:9-98, model.py:92-95,143-146)
// ---------------------------------------------------------------------------------------------------------------
void geo_transpose_pool(const Tensor& geo, Tensor pyramid, int64_t num_levels) {
  check_cuda(geo, "geo"); check_cuda(pyramid, "pyramid");
  TORCH_CHECK(geo.dim() == 5, "geo must be (B, G, D, H, W1), got ", geo.sizes());
  c10::cuda::CUDAGuard guard(geo.device());
  const int64_t B = geo.size(0), G = geo.size(1), D = geo.size(2), H = geo.size(3), W1 = geo.size(4);
  LevelTable t = row_levels(pyramid, B * G * H * W1, D, num_levels, "geo_transpose_pool");
  check_status(nnd_geo_transpose_pool(geo.data_ptr<float>(), B, G, D, H, W1, num_levels, t.ptr.data(), t.pitch.data(),
                                      current_stream(geo)),
               "nnd_geo_transpose_pool");
}

void gev_interleave_pool(const Tensor& src, int64_t layout, int64_t src_pitch, int64_t B, int64_t D, int64_t H, int64_t W1,
                         Tensor interleaved, int64_t num_levels) {
  check_cuda(src, "src"); check_cuda(interleaved, "interleaved");
  same_device(src, interleaved, "gev_interleave_pool");
  c10::cuda::CUDAGuard guard(src.device());
  LevelTable t = interleaved_levels(interleaved, B * H * W1, D, num_levels, "gev_interleave_pool");
  check_status(nnd_gev_interleave_pool(src.data_ptr<float>(), layout, src_pitch, B, 8, D, H, W1, num_levels, t.ptr.data(),
                                       current_stream(src)),
               "nnd_gev_interleave_pool");
}

Tensor gev_lookup(const Tensor& feat_il, const Tensor& geo_il, const Tensor& coords, int64_t D, int64_t num_levels,
                  int64_t radius) {
  check_cuda(feat_il, "feat pyramid"); check_cuda(geo_il, "geo pyramid"); check_cuda(coords, "coords");
  same_device(feat_il, coords, "gev_lookup"); same_device(geo_il, coords, "gev_lookup");
  auto s = nchw(coords, "coords");
  TORCH_CHECK(s[1] == 1, "coords must be (B, 1, H, W), got ", coords.sizes());
  c10::cuda::CUDAGuard guard(coords.device());
  const int64_t pixels = s[0] * s[2] * s[3];
  LevelTable tf = interleaved_levels(feat_il, pixels, D, num_levels, "gev_lookup");
  LevelTable tg = interleaved_levels(geo_il, pixels, D, num_levels, "gev_lookup");
  Tensor out = at::empty({s[0], num_levels * 2 * 8 * (2 * radius + 1), s[2], s[3]}, coords.options());
  check_status(nnd_gev_lookup(tf.ptr.data(), tg.ptr.data(), coords.data_ptr<float>(), s[0], 8, D, s[2], s[3], num_levels, radius,
                              out.data_ptr<float>(), current_stream(coords)),
               "nnd_gev_lookup");
  return out;
}

Tensor soft_argmin(const Tensor& cost) {
  check_cuda(cost, "cost");
  auto s = nchw(cost, "cost");
  c10::cuda::CUDAGuard guard(cost.device());
  Tensor out = at::empty({s[0], 1, s[2], s[3]}, cost.options());
  check_status(nnd_soft_argmin(cost.data_ptr<float>(), s[0], s[1], s[2], s[3], out.data_ptr<float>(), current_stream(cost)),
               "nnd_soft_argmin");
  return out;
}

std::tuple<Tensor, Tensor> gev_squeeze_soft_argmin(const Tensor& geo_level0, const Tensor& weight, const optional<Tensor>& bias,
                                                   int64_t B, int64_t G, int64_t D, int64_t H, int64_t W1, bool return_cost) {
  check_cuda(geo_level0, "geo level 0"); check_cuda(weight, "cv_squeezer.weight");
  same_device(geo_level0, weight, "gev_squeeze_soft_argmin");
  TORCH_CHECK(geo_level0.numel() >= B * H * W1 * D * G, "geo level 0 is smaller than (B, H, W1, D, G)");
  TORCH_CHECK(weight.numel() == G * 27, "cv_squeezer.weight must be (1, ", G, ", 3, 3, 3), got ", weight.sizes());
  const float* bias_ptr = nullptr;
  if (bias.has_value()) {
    check_cuda(*bias, "cv_squeezer.bias");
    bias_ptr = bias->data_ptr<float>();
  }
  c10::cuda::CUDAGuard guard(weight.device());
  Tensor out = at::empty({B, 1, H, W1}, weight.options());
  Tensor cost = return_cost ? at::empty({B, D, H, W1}, weight.options()) : at::empty({0}, weight.options());
  check_status(nnd_gev_squeeze_soft_argmin(geo_level0.data_ptr<float>(), weight.data_ptr<float>(), bias_ptr, B, G, D, H, W1,
                                           out.data_ptr<float>(), return_cost ? cost.data_ptr<float>() : nullptr,
                                           current_stream(weight)),
               "nnd_gev_squeeze_soft_argmin");
  return {out, cost};
}

// ---------------------------------------------------------------------------------------------------------------
// CREStereo: AGCL  (cre_stereo/cost_volume.py:6-154)
// ---------------------------------------------------------------------------------------------------------------
Tensor nchw_to_nhwc(const Tensor& src) {
  check_cuda(src, "src");
  auto s = nchw(src, "src");
  c10::cuda::CUDAGuard guard(src.device());
  Tensor dst = at::empty({s[0], s[2], s[3], s[1]}, src.options());
  check_status(nnd_nchw_to_nhwc(src.data_ptr<float>(), s[0], s[1], s[2], s[3], dst.data_ptr<float>(), current_stream(src)),
               "nnd_nchw_to_nhwc");
  return dst;
}

// fmaps: (N, C, H, W) when nhwc is false, channels-last copies (N, H, W, C) when true
Tensor agcl_offset(const Tensor& f1, const Tensor& f2, const Tensor& flow, const Tensor& extra_offset, bool small_patch,
                   bool nhwc) {
  check_cuda(f1, "fmap1"); check_cuda(f2, "fmap2"); check_cuda(flow, "flow"); check_cuda(extra_offset, "extra_offset");
  same_device(f1, f2, "agcl_offset"); same_device(f1, flow, "agcl_offset"); same_device(f1, extra_offset, "agcl_offset");
  auto sf = nchw(flow, "flow");
  const int64_t N = sf[0], H = sf[2], W = sf[3];
  TORCH_CHECK(f1.dim() == 4 && f1.sizes() == f2.sizes(), "fmap1 and fmap2 must be 4-D of identical shape");
  const int64_t C = nhwc ? f1.size(3) : f1.size(1);
  TORCH_CHECK(sf[1] == 2 && f1.numel() == N * C * H * W, "flow must be (N, 2, H, W) matching the feature maps, got ", flow.sizes());
  TORCH_CHECK(extra_offset.dim() == 4 && extra_offset.size(0) == N && extra_offset.size(1) == 18 && extra_offset.size(2) == H &&
                  extra_offset.size(3) == W,
              "extra_offset must be (N, 18, H, W), got ", extra_offset.sizes());
  c10::cuda::CUDAGuard guard(f1.device());
  Tensor out = at::empty({N, 36, H, W}, flow.options());
  if (nhwc)
    check_status(nnd_agcl_offset_nhwc(f1.data_ptr<float>(), f2.data_ptr<float>(), flow.data_ptr<float>(),
                                      extra_offset.data_ptr<float>(), N, C, H, W, small_patch ? 1 : 0, out.data_ptr<float>(),
                                      current_stream(f1)),
                 "nnd_agcl_offset_nhwc");
  else
    check_status(nnd_agcl_offset(f1.data_ptr<float>(), f2.data_ptr<float>(), flow.data_ptr<float>(), extra_offset.data_ptr<float>(),
                                 N, C, H, W, small_patch ? 1 : 0, out.data_ptr<float>(), current_stream(f1)),
                 "nnd_agcl_offset");
  return out;
}

Tensor agcl_iter(const Tensor& f1, const Tensor& f2, const Tensor& flow, bool small_patch, bool nhwc,
                 const optional<Tensor>& warped_ws) {
  check_cuda(f1, "fmap1"); check_cuda(f2, "fmap2"); check_cuda(flow, "flow");
  same_device(f1, f2, "agcl_iter"); same_device(f1, flow, "agcl_iter");
  auto sf = nchw(flow, "flow");
  const int64_t N = sf[0], H = sf[2], W = sf[3];
  TORCH_CHECK(f1.dim() == 4 && f1.sizes() == f2.sizes(), "fmap1 and fmap2 must be 4-D of identical shape");
  const int64_t C = nhwc ? f1.size(3) : f1.size(1);
  TORCH_CHECK(sf[1] == 2 && f1.numel() == N * C * H * W, "flow must be (N, 2, H, W) matching the feature maps, got ", flow.sizes());
  c10::cuda::CUDAGuard guard(f1.device());
  Tensor out = at::empty({N, 36, H, W}, flow.options());
  if (nhwc) {
    TORCH_CHECK(warped_ws.has_value(), "the channels-last iter kernel needs its (N, H, W, C) workspace");
    check_cuda(*warped_ws, "warped_ws");
    TORCH_CHECK(warped_ws->numel() >= f2.numel(), "warped_ws is smaller than the right feature map");
    check_status(nnd_agcl_iter_nhwc(f1.data_ptr<float>(), f2.data_ptr<float>(), flow.data_ptr<float>(), N, C, H, W,
                                    small_patch ? 1 : 0, warped_ws->data_ptr<float>(), out.data_ptr<float>(), current_stream(f1)),
                 "nnd_agcl_iter_nhwc");
  } else {
    check_status(nnd_agcl_iter(f1.data_ptr<float>(), f2.data_ptr<float>(), flow.data_ptr<float>(), N, C, H, W, small_patch ? 1 : 0,
                               out.data_ptr<float>(), current_stream(f1)),
                 "nnd_agcl_iter");
  }
  return out;
}


// ---- backward of the channels-last AGCL (training) ----
Tensor agcl_warp(const Tensor& f2, const Tensor& flow) {
  check_cuda(f2, "fmap2"); check_cuda(flow, "flow");
  same_device(f2, flow, "agcl_warp");
  auto sf = nchw(flow, "flow");
  TORCH_CHECK(f2.dim() == 4 && f2.size(0) == sf[0] && f2.size(1) == sf[2] && f2.size(2) == sf[3] && sf[1] == 2,
              "agcl_warp: fmap2 must be (N, H, W, C) and flow (N, 2, H, W)");
  c10::cuda::CUDAGuard guard(f2.device());
  Tensor out = at::empty_like(f2);
  check_status(nnd_agcl_warp_nhwc(f2.data_ptr<float>(), flow.data_ptr<float>(), sf[0], f2.size(3), sf[2], sf[3],
                                  out.data_ptr<float>(), current_stream(f2)),
               "nnd_agcl_warp_nhwc");
  return out;
}

// returns (d_fmap1, d_fmap2, d_flow, d_extra), maps channels-last (N, H, W, C)
std::tuple<Tensor, Tensor, Tensor, Tensor> agcl_offset_backward(const Tensor& f1, const Tensor& f2, const Tensor& flow,
                                                                const Tensor& extra_offset, const Tensor& grad_out,
                                                                bool small_patch) {
  check_cuda(f1, "fmap1"); check_cuda(f2, "fmap2"); check_cuda(flow, "flow"); check_cuda(extra_offset, "extra_offset");
  check_cuda(grad_out, "grad_out");
  auto sf = nchw(flow, "flow");
  const int64_t N = sf[0], H = sf[2], W = sf[3];
  TORCH_CHECK(f1.dim() == 4 && f1.sizes() == f2.sizes() && f1.size(0) == N && f1.size(1) == H && f1.size(2) == W && sf[1] == 2,
              "agcl_offset_backward: maps must be (N, H, W, C) matching flow (N, 2, H, W)");
  TORCH_CHECK(grad_out.dim() == 4 && grad_out.size(0) == N && grad_out.size(1) == 36 && grad_out.size(2) == H && grad_out.size(3) == W,
              "agcl_offset_backward: grad_out must be (N, 36, H, W), got ", grad_out.sizes());
  TORCH_CHECK(extra_offset.numel() == N * 18 * H * W, "agcl_offset_backward: extra_offset must be (N, 18, H, W)");
  c10::cuda::CUDAGuard guard(f1.device());
  Tensor d1 = at::empty_like(f1), d2 = at::zeros_like(f2), dflow = at::empty_like(flow), dextra = at::empty_like(extra_offset);
  check_status(nnd_agcl_offset_backward_nhwc(f1.data_ptr<float>(), f2.data_ptr<float>(), flow.data_ptr<float>(),
                                             extra_offset.data_ptr<float>(), grad_out.data_ptr<float>(), N, f1.size(3), H, W,
                                             small_patch ? 1 : 0, d1.data_ptr<float>(), d2.data_ptr<float>(),
                                             dflow.data_ptr<float>(), dextra.data_ptr<float>(), current_stream(f1)),
               "nnd_agcl_offset_backward_nhwc");
  return {d1, d2, dflow, dextra};
}

// returns (d_fmap1, d_fmap2, d_flow); `warped` = agcl_warp(fmap2, flow)
std::tuple<Tensor, Tensor, Tensor> agcl_iter_backward(const Tensor& f1, const Tensor& f2, const Tensor& flow, const Tensor& warped,
                                                      const Tensor& grad_out, bool small_patch, bool left_only) {
  check_cuda(f1, "fmap1"); check_cuda(f2, "fmap2"); check_cuda(flow, "flow"); check_cuda(warped, "warped");
  check_cuda(grad_out, "grad_out");
  auto sf = nchw(flow, "flow");
  const int64_t N = sf[0], H = sf[2], W = sf[3];
  TORCH_CHECK(f1.dim() == 4 && f1.sizes() == f2.sizes() && f1.sizes() == warped.sizes() && f1.size(0) == N && f1.size(1) == H &&
                  f1.size(2) == W && sf[1] == 2,
              "agcl_iter_backward: maps must be (N, H, W, C) matching flow (N, 2, H, W)");
  TORCH_CHECK(grad_out.dim() == 4 && grad_out.size(0) == N && grad_out.size(1) == 36 && grad_out.size(2) == H && grad_out.size(3) == W,
              "agcl_iter_backward: grad_out must be (N, 36, H, W), got ", grad_out.sizes());
  c10::cuda::CUDAGuard guard(f1.device());
  Tensor d1 = at::empty_like(f1);
  if (left_only) {   // the reference detaches the warped right map: only the left features receive a gradient
    check_status(nnd_agcl_iter_backward_nhwc(f1.data_ptr<float>(), f2.data_ptr<float>(), flow.data_ptr<float>(),
                                             warped.data_ptr<float>(), grad_out.data_ptr<float>(), N, f1.size(3), H, W,
                                             small_patch ? 1 : 0, d1.data_ptr<float>(), nullptr, nullptr, nullptr,
                                             current_stream(f1)),
                 "nnd_agcl_iter_backward_nhwc");
    return {d1, at::empty({0}, f1.options()), at::empty({0}, f1.options())};
  }
  Tensor d2 = at::zeros_like(f2), dflow = at::empty_like(flow), dws = at::zeros_like(f2);
  check_status(nnd_agcl_iter_backward_nhwc(f1.data_ptr<float>(), f2.data_ptr<float>(), flow.data_ptr<float>(),
                                           warped.data_ptr<float>(), grad_out.data_ptr<float>(), N, f1.size(3), H, W,
                                           small_patch ? 1 : 0, d1.data_ptr<float>(), d2.data_ptr<float>(),
                                           dflow.data_ptr<float>(), dws.data_ptr<float>(), current_stream(f1)),
               "nnd_agcl_iter_backward_nhwc");
  return {d1, d2, dflow};
}

// ---------------------------------------------------------------------------------------------------------------
// convex upsampling  (raft_stereo/model.py:93-105)
// ---------------------------------------------------------------------------------------------------------------
// mask_layout 0: (N, 9*rate^2, H, W) fp32 NCHW; 1 / 2: the same shape with channels-last strides, fp32 / fp16
Tensor convex_upsample(const Tensor& flow, const Tensor& mask, const optional<Tensor>& mask_bias, int64_t rate, double mask_scale,
                       int64_t mask_layout) {
  check_cuda(flow, "flow");
  auto s = nchw(flow, "flow");
  TORCH_CHECK(s[1] == 1, "flow must be (N, 1, H, W), got ", flow.sizes());
  TORCH_CHECK(mask.is_cuda() && mask.dim() == 4 && mask.size(0) == s[0] && mask.size(1) == 9 * rate * rate && mask.size(2) == s[2] &&
                  mask.size(3) == s[3],
              "mask must be a CUDA (N, 9*rate*rate, H, W) tensor, got ", mask.sizes());
  same_device(flow, mask, "convex_upsample");
  if (mask_layout == 0) {
    check_cuda(mask, "mask");
  } else {
    TORCH_CHECK(mask.scalar_type() == (mask_layout == 2 ? at::kHalf : at::kFloat) && mask.is_contiguous(at::MemoryFormat::ChannelsLast),
                "mask_layout ", mask_layout, " needs a channels-last ", (mask_layout == 2 ? "fp16" : "fp32"), " mask");
  }
  const float* bias_ptr = nullptr;
  if (mask_bias.has_value()) {
    check_cuda(*mask_bias, "mask_bias");
    TORCH_CHECK(mask_bias->numel() == 9 * rate * rate, "mask_bias must have ", 9 * rate * rate, " elements");
    bias_ptr = mask_bias->data_ptr<float>();
  }
  c10::cuda::CUDAGuard guard(flow.device());
  Tensor out = at::empty({s[0], 1, rate * s[2], rate * s[3]}, flow.options());
  check_status(nnd_convex_upsample(flow.data_ptr<float>(), mask.data_ptr(), bias_ptr, s[0], s[2], s[3], rate,
                                   static_cast<float>(mask_scale), mask_layout, out.data_ptr<float>(), current_stream(flow)),
               "nnd_convex_upsample");
  return out;
}

int64_t abi_version() { return nnd_abi_version(); }

}  // namespace

TORCH_LIBRARY(nndepth_b200, m) {
  m.def("abi_version() -> int", &abi_version);
  m.def("corr1d_build(Tensor fmap1, Tensor fmap2, Tensor(a!) pyramid, int num_levels, int precision) -> ()", &corr1d_build);
  m.def("corr1d_build_nhwc_f16(Tensor fmap1, Tensor fmap2, Tensor(a!) pyramid, int num_levels) -> ()", &corr1d_build_nhwc_f16);
  m.def("groupcorr_build(Tensor fmap1, Tensor fmap2, Tensor(a!) pyramid, int num_groups, int group_size, float scale_div, "
        "int num_levels) -> ()", &groupcorr_build);
  m.def("avgpool_pairs(Tensor src, int src_width) -> Tensor", &avgpool_pairs);
  m.def("corr1d_lookup(Tensor pyramid, int width0, Tensor coords, int num_levels, int radius) -> Tensor", &corr1d_lookup);
  m.def("corr1d_lookup_conv1x1(Tensor pyramid, int width0, Tensor coords, int num_levels, int radius, Tensor weight_t, "
        "Tensor? bias, bool relu, int precision, int out_layout) -> Tensor", &corr1d_lookup_conv1x1);
  m.def("corr1d_skew(Tensor pyramid, int B, int H, int W1, int W2, int num_levels) -> Tensor", &corr1d_skew);
  m.def("corr1d_lookup_skewed(Tensor skewed, int width0, Tensor coords, int num_levels, int radius) -> Tensor", &corr1d_lookup_skewed);
  m.def("corr1d_lookup_conv1x1_skewed(Tensor skewed, int width0, Tensor coords, int num_levels, int radius, Tensor weight_t, "
        "Tensor? bias, bool relu, int out_layout) -> Tensor", &corr1d_lookup_conv1x1_skewed);
  m.def("corr1d_lookup_backward(Tensor grad_out, Tensor coords, int width0, int num_levels, int radius) -> Tensor",
        &corr1d_lookup_backward);
  m.def("pyramid_unpool_(Tensor(a!) d_pyramid, int rows, int width0, int num_levels) -> ()", &pyramid_unpool_);
  m.def("volume_grad(Tensor d_pyramid, Tensor f_other, int W1, int W2, int num_groups, int group_size, float scale_div, int which) "
        "-> Tensor", &volume_grad);
  m.def("corr1d_lookup_indices(int[] widths, Tensor coords, int num_levels, int radius) -> (Tensor, Tensor)",
        &corr1d_lookup_indices);
  m.def("group_lookup(Tensor pyramid_a, Tensor? pyramid_b, int width0, Tensor coords, int num_groups, int num_levels, int radius, "
        "int mode) -> Tensor", &group_lookup);
  m.def("geo_transpose_pool(Tensor geo, Tensor(a!) pyramid, int num_levels) -> ()", &geo_transpose_pool);
  m.def("gev_interleave_pool(Tensor src, int layout, int src_pitch, int B, int D, int H, int W1, Tensor(a!) interleaved, "
        "int num_levels) -> ()", &gev_interleave_pool);
  m.def("gev_lookup(Tensor feat, Tensor geo, Tensor coords, int D, int num_levels, int radius) -> Tensor", &gev_lookup);
  m.def("soft_argmin(Tensor cost) -> Tensor", &soft_argmin);
  m.def("gev_squeeze_soft_argmin(Tensor geo_level0, Tensor weight, Tensor? bias, int B, int G, int D, int H, int W1, "
        "bool return_cost) -> (Tensor, Tensor)", &gev_squeeze_soft_argmin);
  m.def("nchw_to_nhwc(Tensor src) -> Tensor", &nchw_to_nhwc);
  m.def("agcl_offset(Tensor fmap1, Tensor fmap2, Tensor flow, Tensor extra_offset, bool small_patch, bool nhwc) -> Tensor",
        &agcl_offset);
  m.def("agcl_iter(Tensor fmap1, Tensor fmap2, Tensor flow, bool small_patch, bool nhwc, Tensor? warped_ws) -> Tensor", &agcl_iter);
  m.def("agcl_warp(Tensor fmap2, Tensor flow) -> Tensor", &agcl_warp);
  m.def("agcl_offset_backward(Tensor fmap1, Tensor fmap2, Tensor flow, Tensor extra_offset, Tensor grad_out, bool small_patch) "
        "-> (Tensor, Tensor, Tensor, Tensor)", &agcl_offset_backward);
  m.def("agcl_iter_backward(Tensor fmap1, Tensor fmap2, Tensor flow, Tensor warped, Tensor grad_out, bool small_patch, "
        "bool left_only) -> (Tensor, Tensor, Tensor)", &agcl_iter_backward);
  m.def("convex_upsample(Tensor flow, Tensor mask, Tensor? mask_bias, int rate, float mask_scale, int mask_layout) -> Tensor",
        &convex_upsample);
}

// Letterbox preprocessing on the device: the reference's resize_image (utils/utils.py:16-30) -- Pillow
// Image.resize(..., Image.BICUBIC) pasted on a grey canvas -- producing the uint8 NHWC batch the stem kernel reads
// (SURVEY.md section 8 row f2).  Pillow's 8-bit resampler is integer arithmetic (22-bit fixed-point weights, a
// horizontal pass rounded to uint8, then a vertical pass), so the result is reproduced BIT FOR BIT:
//   * tod_resample_coeffs_bicubic (host, double arithmetic in Pillow's order of operations) builds the per-output-index
//     windows and fixed-point weights of one axis: Resample.c precompute_coeffs + normalize_coeffs_8bpc;
//   * letterbox_h_kernel / letterbox_v_kernel are the two passes: ImagingResampleHorizontal_8bpc /
//     ImagingResampleVertical_8bpc; the vertical pass writes the whole canvas (pad value outside the pasted image),
//     so there is no separate fill.
// Bandwidth-trivial next to the network (a 640x480 source is 0.9 MB in, 1.2 MB out); one thread per output pixel,
// windows of neighbouring threads overlap in L1.
#include <cmath>

#include "tod_common.cuh"

namespace tod {

constexpr int kPrecisionBits = 32 - 8 - 2;   // Resample.c PRECISION_BITS

static double bicubic_filter(double x) {     // Resample.c bicubic_filter, a = -0.5
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kPrecisionBits;                       // arithmetic shift, like Pillow's lookup index
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

struct LbParams {
  const uint8_t* src;
  uint8_t* tmp;
  uint8_t* dst;
  long long src_image_stride, dst_image_stride;
  int n, src_h, src_w, dst_h, dst_w, new_h, new_w, off_y, off_x, pad;
  const int* xb;
  const int* xk;
  int xks;
  const int* yb;
  const int* yk;
  int yks;
};

// horizontal pass: (n, src_h, src_w, 3) -> tmp (n, src_h, new_w, 3)
__global__ void __launch_bounds__(256) letterbox_h_kernel(const LbParams p, int canvas) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.src_h * p.new_w) return;
  const int img = blockIdx.y;
  const int y = idx / p.new_w, xx = idx - y * p.new_w;
  const int xmin = __ldg(p.xb + 2 * xx), cnt = __ldg(p.xb + 2 * xx + 1);
  const int* k = p.xk + static_cast<long long>(xx) * p.xks;
  const uint8_t* row = p.src + img * p.src_image_stride + (static_cast<long long>(y) * p.src_w + xmin) * 3;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  for (int x = 0; x < cnt; ++x) {
    const int w = __ldg(k + x);
    s0 += static_cast<int>(row[3 * x]) * w;
    s1 += static_cast<int>(row[3 * x + 1]) * w;
    s2 += static_cast<int>(row[3 * x + 2]) * w;
  }
  // canvas: the intermediate has canvas-width rows with the resized row at off_x (word-wise vertical pass)
  uint8_t* o = canvas ? p.tmp + ((static_cast<long long>(img) * p.src_h + y) * p.dst_w + p.off_x + xx) * 3
                      : p.tmp + ((static_cast<long long>(img) * p.src_h + y) * p.new_w + xx) * 3;
  o[0] = clip8(s0);
  o[1] = clip8(s1);
  o[2] = clip8(s2);
}

// vertical pass + paste: `in` is tmp (or the source when the width is unchanged), rows of new_w pixels
__global__ void __launch_bounds__(256) letterbox_v_kernel(const LbParams p, const uint8_t* in, long long in_image_stride,
                                                          int need_v) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.dst_h * p.dst_w) return;
  const int img = blockIdx.y;
  const int y = idx / p.dst_w, x = idx - y * p.dst_w;
  uint8_t* o = p.dst + img * p.dst_image_stride + static_cast<long long>(idx) * 3;
  const int yy = y - p.off_y, xx = x - p.off_x;
  if (yy < 0 || yy >= p.new_h || xx < 0 || xx >= p.new_w) {
    o[0] = o[1] = o[2] = static_cast<uint8_t>(p.pad);
    return;
  }
  const uint8_t* base = in + img * in_image_stride + static_cast<long long>(xx) * 3;
  const long long pitch = static_cast<long long>(p.new_w) * 3;
  if (!need_v) {
    const uint8_t* s = base + yy * pitch;
    o[0] = s[0];
    o[1] = s[1];
    o[2] = s[2];
    return;
  }
  const int ymin = __ldg(p.yb + 2 * yy), cnt = __ldg(p.yb + 2 * yy + 1);
  const int* k = p.yk + static_cast<long long>(yy) * p.yks;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  const uint8_t* s = base + ymin * pitch;
  for (int j = 0; j < cnt; ++j, s += pitch) {
    const int w = __ldg(k + j);
    s0 += static_cast<int>(s[0]) * w;
    s1 += static_cast<int>(s[1]) * w;
    s2 += static_cast<int>(s[2]) * w;
  }
  o[0] = clip8(s0);
  o[1] = clip8(s1);
  o[2] = clip8(s2);
}

// ---- fast path (canvas rows of a multiple of 4 bytes).  The intermediate image is kept in CANVAS byte coordinates
// (rows of dst_w * 3 bytes, the resized row starting at off_x * 3), so that the vertical pass -- which is channel
// agnostic: output byte b of a row depends on byte b of the rows above and below only -- runs on aligned 32-bit words.
// (Measured and dropped: staging the source span + weights of 128 output pixels x 16 rows in shared memory for the
// horizontal pass was SLOWER than one thread per output pixel through L1: 985 vs 758 us for 64 x 1080p.)
// vertical pass + paste on 4-byte words of the canvas row
__global__ void __launch_bounds__(256) letterbox_v_word_kernel(const LbParams p, long long tmp_image_stride, int need_v) {
  const int words_per_row = p.dst_w * 3 / 4;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.dst_h * words_per_row) return;
  const int img = blockIdx.y;
  const int y = idx / words_per_row, wd = idx - y * words_per_row;
  uint32_t* o = reinterpret_cast<uint32_t*>(p.dst + img * p.dst_image_stride) + idx;
  const uint32_t pad4 = static_cast<uint32_t>(p.pad & 0xff) * 0x01010101u;
  const int yy = y - p.off_y;
  const int b0 = wd * 4, lo = p.off_x * 3, hi = (p.off_x + p.new_w) * 3;      // byte range of the image inside a row
  if (yy < 0 || yy >= p.new_h || b0 + 4 <= lo || b0 >= hi) {
    *o = pad4;
    return;
  }
  const long long pitch = static_cast<long long>(p.dst_w) * 3;
  const uint8_t* base = p.tmp + img * tmp_image_stride + b0;
  uint32_t r;
  if (!need_v) {
    r = *reinterpret_cast<const uint32_t*>(base + yy * pitch);
  } else {
    const int ymin = __ldg(p.yb + 2 * yy), cnt = __ldg(p.yb + 2 * yy + 1);
    const int* k = p.yk + static_cast<long long>(yy) * p.yks;
    const int half = 1 << (kPrecisionBits - 1);
    int a0 = half, a1 = half, a2 = half, a3 = half;
    const uint8_t* s = base + ymin * pitch;
    for (int j = 0; j < cnt; ++j, s += pitch) {
      const uint32_t v = *reinterpret_cast<const uint32_t*>(s);
      const int w = __ldg(k + j);
      a0 += static_cast<int>(v & 0xffu) * w;
      a1 += static_cast<int>((v >> 8) & 0xffu) * w;
      a2 += static_cast<int>((v >> 16) & 0xffu) * w;
      a3 += static_cast<int>(v >> 24) * w;
    }
    r = static_cast<uint32_t>(clip8(a0)) | (static_cast<uint32_t>(clip8(a1)) << 8) | (static_cast<uint32_t>(clip8(a2)) << 16) |
        (static_cast<uint32_t>(clip8(a3)) << 24);
  }
  if (b0 < lo || b0 + 4 > hi) {   // word straddles the image edge: pad the bytes outside
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (b0 + i < lo || b0 + i >= hi) r = (r & ~(0xffu << (8 * i))) | ((pad4 & 0xffu) << (8 * i));
  }
  *o = r;
}

}  // namespace tod

using namespace tod;

extern "C" int tod_resample_coeffs_bicubic(int32_t in_size, int32_t out_size, int32_t* h_bounds, int32_t* h_coef,
                                           int32_t* ksize_out) {
  TOD_CHECK_ARG(in_size > 0 && out_size > 0 && ksize_out != nullptr, "resample coeffs: bad sizes %d -> %d", in_size, out_size);
  // float in0 = 0, in1 = in_size in Pillow's signature: (double)(in1 - in0) / outSize
  const double scale = static_cast<double>(static_cast<float>(in_size) - 0.0f) / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 2.0 * filterscale;
  const int ksize = static_cast<int>(std::ceil(support)) * 2 + 1;
  *ksize_out = ksize;
  if (h_bounds == nullptr || h_coef == nullptr) return TOD_OK;   // size query
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = 0.0 + (xx + 0.5) * scale;
    double ww = 0.0;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    int32_t* k = h_coef + static_cast<long long>(xx) * ksize;
    double kd[64 * 4];
    TOD_CHECK_ARG(xmax <= 256, "resample coeffs: window of %d taps (scale %.1f) exceeds the supported 256", xmax, scale);
    for (int x = 0; x < xmax; ++x) {
      const double w = bicubic_filter((x + xmin - center + 0.5) * ss);
      kd[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x) {
      if (ww != 0.0) kd[x] /= ww;
      k[x] = kd[x] < 0 ? static_cast<int>(-0.5 + kd[x] * (1 << kPrecisionBits)) : static_cast<int>(0.5 + kd[x] * (1 << kPrecisionBits));
    }
    for (int x = xmax; x < ksize; ++x) k[x] = 0;
    h_bounds[2 * xx] = xmin;
    h_bounds[2 * xx + 1] = xmax;
  }
  return TOD_OK;
}

extern "C" int tod_letterbox_bicubic_u8(const tod_letterbox_desc* d, void* stream) {
  TOD_CHECK_ARG(d != nullptr && d->d_src != nullptr && d->d_dst != nullptr, "letterbox: null pointer");
  TOD_CHECK_ARG(d->n > 0 && d->n <= 65535 && d->src_h > 0 && d->src_w > 0 && d->dst_h > 0 && d->dst_w > 0, "letterbox: bad sizes");
  TOD_CHECK_ARG(d->new_h > 0 && d->new_w > 0 && d->off_y >= 0 && d->off_x >= 0 && d->off_y + d->new_h <= d->dst_h &&
                    d->off_x + d->new_w <= d->dst_w,
                "letterbox: resized image %dx%d at (%d, %d) does not fit the %dx%d canvas", d->new_h, d->new_w, d->off_y,
                d->off_x, d->dst_h, d->dst_w);
  const bool need_h = d->new_w != d->src_w, need_v = d->new_h != d->src_h;
  TOD_CHECK_ARG(!need_h || (d->d_xbounds != nullptr && d->d_xcoef != nullptr && d->xksize > 0 && d->d_tmp != nullptr),
                "letterbox: horizontal pass needs bounds, coefficients and the intermediate buffer");
  TOD_CHECK_ARG(!need_v || (d->d_ybounds != nullptr && d->d_ycoef != nullptr && d->yksize > 0),
                "letterbox: vertical pass needs bounds and coefficients");
  TOD_CHECK_ARG(d->src_image_stride >= static_cast<int64_t>(d->src_h) * d->src_w * 3 &&
                    d->dst_image_stride >= static_cast<int64_t>(d->dst_h) * d->dst_w * 3,
                "letterbox: image strides smaller than an image");
  TOD_CHECK_ARG(static_cast<long long>(d->src_h) * d->new_w < (1ll << 31) && static_cast<long long>(d->dst_h) * d->dst_w < (1ll << 31),
                "letterbox: image too large");
  LbParams p;
  p.src = d->d_src;
  p.tmp = d->d_tmp;
  p.dst = d->d_dst;
  p.src_image_stride = d->src_image_stride;
  p.dst_image_stride = d->dst_image_stride;
  p.n = d->n;
  p.src_h = d->src_h;
  p.src_w = d->src_w;
  p.dst_h = d->dst_h;
  p.dst_w = d->dst_w;
  p.new_h = d->new_h;
  p.new_w = d->new_w;
  p.off_y = d->off_y;
  p.off_x = d->off_x;
  p.pad = d->pad_value;
  p.xb = d->d_xbounds;
  p.xk = d->d_xcoef;
  p.xks = d->xksize;
  p.yb = d->d_ybounds;
  p.yk = d->d_ycoef;
  p.yks = d->yksize;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  // fast path: the intermediate in canvas byte coordinates, vertical pass on aligned 32-bit words
  if (need_h && (d->dst_w * 3) % 4 == 0 && (reinterpret_cast<uintptr_t>(d->d_tmp) & 3) == 0 &&
      (reinterpret_cast<uintptr_t>(d->d_dst) & 3) == 0 && d->dst_image_stride % 4 == 0) {
    const long long tmp_stride = static_cast<long long>(d->src_h) * d->dst_w * 3;
    const unsigned hblocks = static_cast<unsigned>((static_cast<long long>(d->src_h) * d->new_w + 255) / 256);
    letterbox_h_kernel<<<dim3(hblocks, d->n), 256, 0, st>>>(p, 1);
    if ((rc = check_cuda(cudaGetLastError(), "letterbox_h_kernel launch")) != TOD_OK) return rc;
    const long long words = static_cast<long long>(d->dst_h) * (d->dst_w * 3 / 4);
    letterbox_v_word_kernel<<<dim3(static_cast<unsigned>((words + 255) / 256), d->n), 256, 0, st>>>(p, tmp_stride, need_v ? 1 : 0);
    return check_cuda(cudaGetLastError(), "letterbox_v_word_kernel launch");
  }
  if (need_h) {
    const unsigned blocks = static_cast<unsigned>((static_cast<long long>(d->src_h) * d->new_w + 255) / 256);
    letterbox_h_kernel<<<dim3(blocks, d->n), 256, 0, st>>>(p, 0);
    if ((rc = check_cuda(cudaGetLastError(), "letterbox_h_kernel launch")) != TOD_OK) return rc;
  }
  const uint8_t* in = need_h ? d->d_tmp : d->d_src;
  const long long in_stride = need_h ? static_cast<long long>(d->src_h) * d->new_w * 3 : d->src_image_stride;
  const unsigned blocks = static_cast<unsigned>((static_cast<long long>(d->dst_h) * d->dst_w + 255) / 256);
  letterbox_v_kernel<<<dim3(blocks, d->n), 256, 0, st>>>(p, in, in_stride, need_v ? 1 : 0);
  return check_cuda(cudaGetLastError(), "letterbox_v_kernel launch");
}

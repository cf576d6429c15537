// Stem: Conv2d(3, cout, 3, s2, p1) + folded BN + SiLU, NCHW fp32 in -> NHWC bf16 out.
// Replaces backbone.stem (model/backbone.py:20; Conv.forward model/blocks.py:52-54).
// K = 27 is too thin for a tensor-core tile and the layer is HBM-bound (reads the fp32 image once, writes
// cout bf16 per output pixel), so it runs on CUDA cores: one thread per output pixel, all cout channels.
// The folded weights travel as a __grid_constant__ kernel parameter: with fully unrolled loops every FFMA takes
// its weight straight from the constant bank (no shared/global weight loads at all); 16-byte vector stores.
#include <cstring>

#include "tod_common.cuh"

namespace tod {

constexpr int kStemThreads = 128;

template <int COUT>
struct StemWeights {
  float w[COUT * 27];   // [cout][c*9 + kh*3 + kw]
  float b[COUT];
};

template <int COUT>
__global__ void __launch_bounds__(kStemThreads) stem_conv_kernel(const float* __restrict__ x,
                                                                 __nv_bfloat16* __restrict__ out, int hin, int win,
                                                                 int out_pitch,
                                                                 const __grid_constant__ StemWeights<COUT> wt) {
  const int hout = hin >> 1, wout = win >> 1;
  const int pix = blockIdx.x * kStemThreads + threadIdx.x;
  const int n = blockIdx.y;
  if (pix >= hout * wout) return;
  const int oh = pix / wout, ow = pix - oh * wout;

  float v[27];
  const size_t plane = static_cast<size_t>(hin) * win;
  const float* xn = x + static_cast<size_t>(n) * 3 * plane;
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = 2 * oh + kh - 1;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int iw = 2 * ow + kw - 1;
        const bool ok = ih >= 0 && ih < hin && iw >= 0 && iw < win;
        v[c * 9 + kh * 3 + kw] = ok ? __ldg(xn + c * plane + static_cast<size_t>(ih) * win + iw) : 0.f;
      }
    }

  __nv_bfloat16* o = out + (static_cast<size_t>(n) * hout * wout + pix) * out_pitch;
#pragma unroll
  for (int co = 0; co < COUT; co += 8) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = wt.b[co + j];
#pragma unroll
      for (int k = 0; k < 27; ++k) a = fmaf(v[k], wt.w[(co + j) * 27 + k], a);
      acc[j] = silu_f(a);
    }
    uint4 ov;
    ov.x = pack_bf16x2(acc[0], acc[1]);
    ov.y = pack_bf16x2(acc[2], acc[3]);
    ov.z = pack_bf16x2(acc[4], acc[5]);
    ov.w = pack_bf16x2(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(o + co) = ov;
  }
}

template <int COUT>
static int launch_stem(const float* d_x, const float* h_w, const float* h_bias, void* d_out, int batch, int hin, int win,
                       int out_pitch, cudaStream_t st) {
  StemWeights<COUT> wt;
  memcpy(wt.w, h_w, sizeof(wt.w));
  if (h_bias) memcpy(wt.b, h_bias, sizeof(wt.b)); else memset(wt.b, 0, sizeof(wt.b));
  dim3 grid(ceil_div((hin / 2) * (win / 2), kStemThreads), batch, 1);
  stem_conv_kernel<COUT><<<grid, kStemThreads, 0, st>>>(d_x, reinterpret_cast<__nv_bfloat16*>(d_out), hin, win, out_pitch, wt);
  TOD_CHECK_LAUNCH("stem_conv_kernel launch");
  return TOD_OK;
}

}  // namespace tod

using namespace tod;

extern "C" int tod_stem_conv_nchw_f32(const float* d_x, const float* h_w, const float* h_bias, void* d_out,
                                      int32_t batch, int32_t hin, int32_t win, int32_t cout, int32_t out_pitch,
                                      void* stream) {
  TOD_CHECK_ARG(d_x && h_w && d_out, "stem: null pointer");
  TOD_CHECK_ARG(batch > 0 && hin > 0 && win > 0 && hin % 2 == 0 && win % 2 == 0, "stem: bad shape %d x %d x %d", batch,
                hin, win);
  TOD_CHECK_ARG(out_pitch >= cout && out_pitch % 8 == 0, "stem: out_pitch %d", out_pitch);
  TOD_CHECK_ARG((reinterpret_cast<uintptr_t>(d_out) & 15) == 0, "stem: output must be 16-byte aligned");
  TOD_CHECK_ARG(batch <= 65535, "stem: batch too large");
  auto st = static_cast<cudaStream_t>(stream);
  switch (cout) {
    case 16: return launch_stem<16>(d_x, h_w, h_bias, d_out, batch, hin, win, out_pitch, st);
    case 32: return launch_stem<32>(d_x, h_w, h_bias, d_out, batch, hin, win, out_pitch, st);
    case 48: return launch_stem<48>(d_x, h_w, h_bias, d_out, batch, hin, win, out_pitch, st);
    case 64: return launch_stem<64>(d_x, h_w, h_bias, d_out, batch, hin, win, out_pitch, st);
    case 96: return launch_stem<96>(d_x, h_w, h_bias, d_out, batch, hin, win, out_pitch, st);
    case 128: return launch_stem<128>(d_x, h_w, h_bias, d_out, batch, hin, win, out_pitch, st);
    default:
      set_error("stem: unsupported cout %d (supported: 16, 32, 48, 64, 96, 128)", cout);
      return TOD_ERR_UNSUPPORTED;
  }
}

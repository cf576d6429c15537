// Stem: Conv2d(3, cout, 3, s2, p1) + folded BN + SiLU, NCHW fp32 in -> NHWC bf16 out.
// Replaces backbone.stem (model/backbone.py:20; Conv.forward model/blocks.py:52-54).
// K = 27 is too thin for a tensor-core tile and the layer is HBM-bound (reads the fp32 image once, writes
// cout bf16 per output pixel), so it runs on CUDA cores: one thread per output pixel, weights broadcast
// from shared memory, 16-byte vector stores.
#include "tod_common.cuh"

namespace tod {

constexpr int kStemThreads = 128;
constexpr int kStemMaxCout = 128;

__global__ void __launch_bounds__(kStemThreads) stem_conv_kernel(const float* __restrict__ x,
                                                                 const float* __restrict__ w,
                                                                 const float* __restrict__ bias,
                                                                 __nv_bfloat16* __restrict__ out, int batch, int hin,
                                                                 int win, int cout, int out_pitch) {
  __shared__ float sw[kStemMaxCout * 27];
  __shared__ float sb[kStemMaxCout];
  for (int i = threadIdx.x; i < cout * 27; i += kStemThreads) sw[i] = w[i];
  for (int i = threadIdx.x; i < cout; i += kStemThreads) sb[i] = bias ? bias[i] : 0.f;
  __syncthreads();

  const int hout = hin >> 1, wout = win >> 1;
  const int ow = blockIdx.x * kStemThreads + threadIdx.x;
  const int oh = blockIdx.y;
  const int n = blockIdx.z;
  if (ow >= wout) return;

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

  __nv_bfloat16* o = out + ((static_cast<size_t>(n) * hout + oh) * wout + ow) * out_pitch;
  for (int co = 0; co < cout; co += 8) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = sb[co + j];
      const float* wr = sw + (co + j) * 27;
#pragma unroll
      for (int k = 0; k < 27; ++k) a = fmaf(v[k], wr[k], a);
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

}  // namespace tod

using namespace tod;

extern "C" int tod_stem_conv_nchw_f32(const float* d_x, const float* d_w, const float* d_bias, void* d_out,
                                      int32_t batch, int32_t hin, int32_t win, int32_t cout, int32_t out_pitch,
                                      void* stream) {
  TOD_CHECK_ARG(d_x && d_w && d_out, "stem: null pointer");
  TOD_CHECK_ARG(batch > 0 && hin > 0 && win > 0 && hin % 2 == 0 && win % 2 == 0, "stem: bad shape %d x %d x %d", batch,
                hin, win);
  TOD_CHECK_ARG(cout > 0 && cout % 8 == 0 && cout <= kStemMaxCout, "stem: cout %d must be a multiple of 8, <= %d", cout,
                kStemMaxCout);
  TOD_CHECK_ARG(out_pitch >= cout && out_pitch % 8 == 0, "stem: out_pitch %d", out_pitch);
  TOD_CHECK_ARG((reinterpret_cast<uintptr_t>(d_out) & 15) == 0, "stem: output must be 16-byte aligned");
  TOD_CHECK_ARG(hin / 2 <= 65535 && batch <= 65535, "stem: grid too large");
  dim3 grid(ceil_div(win / 2, kStemThreads), hin / 2, batch);
  stem_conv_kernel<<<grid, kStemThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      d_x, d_w, d_bias, reinterpret_cast<__nv_bfloat16*>(d_out), batch, hin, win, cout, out_pitch);
  TOD_CHECK_LAUNCH("stem_conv_kernel launch");
  return TOD_OK;
}

// CBAM (reference model/blocks.py:190-223) on NHWC bf16 activations -- SURVEY.md section 8 row f1.
//   channel attention  ca[n,c] = sigmoid(fc2(relu(fc1(avgpool(x)))) + fc2(relu(fc1(maxpool(x)))))      (:208-210)
//   x1 = x * ca                                                                                         (:213)
//   spatial attention  sa[n,p] = sigmoid(conv_kxk(cat[mean_c(x1), max_c(x1)]))                           (:216-218)
//   y = x1 * sa                                                                                         (:221)
// Four bandwidth-bound passes (x is read three times and written once; everything else is tiny and f32):
//   cbam_pool_kernel   per (image, pixel chunk): per-channel sum / max partials              -> work.partial
//   cbam_mlp_kernel    per image: reduce the partials, the two-layer MLP, sigmoid             -> work.ca
//   cbam_stats_kernel  per pixel: mean / max over channels of x * ca                          -> work.stats
//   cbam_apply_kernel  per pixel: k x k conv over the stats (zero padding), sigmoid, y = x * ca * sa
// The reduction order of the pooled mean is fixed (chunk partials summed in order): results are deterministic.
#include "tod_common.cuh"

namespace tod {

constexpr int kCbamThreads = 256;
constexpr int kCbamChunkPx = 512;    // pixels per pooling block (2048 left a 6400-pixel map with 4 x batch blocks: under-filled)

struct CbamParams {
  const __nv_bfloat16* x;
  __nv_bfloat16* out;
  const float* fc1;
  const float* fc2;
  const float* conv;
  float* partial;   // [B][chunks][2][C]
  float* ca;        // [B][C]
  float* stats;     // [B][HW][2]
  float* sa;        // [B][HW] spatial scale (vector-mapped path)
  int batch, hw, h, w, c, hidden, ksize, x_pitch, out_pitch, chunks;
};

__device__ __forceinline__ void bf16x8_to_f32(const uint4 v, float (&f)[8]) {
  const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(u[i] << 16);
    f[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
  }
}

__global__ void __launch_bounds__(kCbamThreads) cbam_pool_kernel(const CbamParams p) {
  extern __shared__ float red[];   // [R][V][16]: 8 sums then 8 maxima
  const int V = p.c >> 3, R = kCbamThreads / V;
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int r = threadIdx.x / V, v = threadIdx.x - r * V;
  const int p0 = chunk * kCbamChunkPx, p1 = min(p0 + kCbamChunkPx, p.hw);
  float s[8], m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    s[i] = 0.0f;
    m[i] = -INFINITY;
  }
  if (r < R) {
    const __nv_bfloat16* base = p.x + static_cast<size_t>(n) * p.hw * p.x_pitch + v * 8;
    for (int px = p0 + r; px < p1; px += R) {
      float f[8];
      bf16x8_to_f32(*reinterpret_cast<const uint4*>(base + static_cast<size_t>(px) * p.x_pitch), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i] += f[i];
        m[i] = fmaxf(m[i], f[i]);
      }
    }
    float* o = red + (r * V + v) * 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o[i] = s[i];
      o[8 + i] = m[i];
    }
  }
  __syncthreads();
  if (r == 0) {   // fixed-order reduction over the R rows
    for (int rr = 1; rr < R; ++rr) {
      const float* o = red + (rr * V + v) * 16;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i] += o[i];
        m[i] = fmaxf(m[i], o[8 + i]);
      }
    }
    float* dst = p.partial + (static_cast<size_t>(n) * p.chunks + chunk) * 2 * p.c + v * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dst[i] = s[i];
      dst[p.c + i] = m[i];
    }
  }
}

__global__ void __launch_bounds__(kCbamThreads) cbam_mlp_kernel(const CbamParams p) {
  extern __shared__ float sm[];   // avg[C], max[C], hid[2 * hidden]
  float* avg = sm;
  float* mx = sm + p.c;
  float* hid = sm + 2 * p.c;
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < p.c; c += kCbamThreads) {
    float s = 0.0f, m = -INFINITY;
    for (int k = 0; k < p.chunks; ++k) {
      const float* src = p.partial + (static_cast<size_t>(n) * p.chunks + k) * 2 * p.c;
      s += src[c];
      m = fmaxf(m, src[p.c + c]);
    }
    avg[c] = s / static_cast<float>(p.hw);
    mx[c] = m;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < 2 * p.hidden; j += kCbamThreads) {
    const float* w = p.fc1 + static_cast<size_t>(j % p.hidden) * p.c;
    const float* in = j < p.hidden ? avg : mx;
    float a = 0.0f;
    for (int c = 0; c < p.c; ++c) a = fmaf(w[c], in[c], a);
    hid[j] = fmaxf(a, 0.0f);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < p.c; c += kCbamThreads) {
    const float* w = p.fc2 + static_cast<size_t>(c) * p.hidden;
    float a = 0.0f, b = 0.0f;
    for (int j = 0; j < p.hidden; ++j) {
      a = fmaf(w[j], hid[j], a);
      b = fmaf(w[j], hid[p.hidden + j], b);
    }
    p.ca[static_cast<size_t>(n) * p.c + c] = 1.0f / (1.0f + __expf(-(a + b)));
  }
}

__global__ void __launch_bounds__(kCbamThreads) cbam_stats_kernel(const CbamParams p) {
  const long long idx = static_cast<long long>(blockIdx.x) * kCbamThreads + threadIdx.x;
  if (idx >= static_cast<long long>(p.batch) * p.hw) return;
  const int n = static_cast<int>(idx / p.hw);
  const __nv_bfloat16* row = p.x + static_cast<size_t>(idx) * p.x_pitch;
  const float* ca = p.ca + static_cast<size_t>(n) * p.c;
  float s = 0.0f, m = -INFINITY;
  for (int v = 0; v < p.c; v += 8) {
    float f[8];
    bf16x8_to_f32(*reinterpret_cast<const uint4*>(row + v), f);
    const float4 c0 = __ldg(reinterpret_cast<const float4*>(ca + v)), c1 = __ldg(reinterpret_cast<const float4*>(ca + v + 4));
    const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float t = f[i] * cc[i];
      s += t;
      m = fmaxf(m, t);
    }
  }
  reinterpret_cast<float2*>(p.stats)[idx] = make_float2(s / static_cast<float>(p.c), m);
}

__global__ void __launch_bounds__(kCbamThreads) cbam_apply_kernel(const CbamParams p) {
  const long long idx = static_cast<long long>(blockIdx.x) * kCbamThreads + threadIdx.x;
  if (idx >= static_cast<long long>(p.batch) * p.hw) return;
  const int n = static_cast<int>(idx / p.hw);
  const int pix = static_cast<int>(idx - static_cast<long long>(n) * p.hw);
  const int y = pix / p.w, x = pix - y * p.w;
  const int k = p.ksize, pad = k >> 1;
  const float2* st = reinterpret_cast<const float2*>(p.stats) + static_cast<size_t>(n) * p.hw;
  float a = 0.0f;
  for (int dy = 0; dy < k; ++dy) {
    const int yy = y + dy - pad;
    if (yy < 0 || yy >= p.h) continue;
    for (int dx = 0; dx < k; ++dx) {
      const int xx = x + dx - pad;
      if (xx < 0 || xx >= p.w) continue;
      const float2 s = __ldg(st + yy * p.w + xx);
      a = fmaf(__ldg(p.conv + dy * k + dx), s.x, a);
      a = fmaf(__ldg(p.conv + k * k + dy * k + dx), s.y, a);
    }
  }
  const float sa = 1.0f / (1.0f + __expf(-a));
  const __nv_bfloat16* row = p.x + static_cast<size_t>(idx) * p.x_pitch;
  __nv_bfloat16* orow = p.out + static_cast<size_t>(idx) * p.out_pitch;
  const float* ca = p.ca + static_cast<size_t>(n) * p.c;
  for (int v = 0; v < p.c; v += 8) {
    float f[8];
    bf16x8_to_f32(*reinterpret_cast<const uint4*>(row + v), f);
    const float4 c0 = __ldg(reinterpret_cast<const float4*>(ca + v)), c1 = __ldg(reinterpret_cast<const float4*>(ca + v + 4));
    const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
    uint4 o;
    o.x = pack_bf16x2(f[0] * cc[0] * sa, f[1] * cc[1] * sa);
    o.y = pack_bf16x2(f[2] * cc[2] * sa, f[3] * cc[3] * sa);
    o.z = pack_bf16x2(f[4] * cc[4] * sa, f[5] * cc[5] * sa);
    o.w = pack_bf16x2(f[6] * cc[6] * sa, f[7] * cc[7] * sa);
    *reinterpret_cast<uint4*>(orow + v) = o;
  }
}

// ---- vector-mapped passes (C / 8 a power of two <= 32): one thread per (pixel, 8-channel vector), so a warp reads
// 512 contiguous bytes; the per-pixel reduction over the V lanes of a pixel is a shuffle butterfly.  The thread-per-pixel
// kernels above streamed one 16-byte piece of 32 different rows per load instruction (2.4-2.6 TB/s for the whole block).
__global__ void __launch_bounds__(kCbamThreads) cbam_stats_vec_kernel(const CbamParams p, int V, int vshift) {
  const long long e = static_cast<long long>(blockIdx.x) * kCbamThreads + threadIdx.x;
  const long long total = static_cast<long long>(p.batch) * p.hw * V;
  const bool on = e < total;
  const long long pixel = on ? (e >> vshift) : 0;
  const int v = static_cast<int>(e & (V - 1));
  const int n = static_cast<int>(pixel / p.hw);
  float s = 0.0f, m = -INFINITY;
  if (on) {
    float f[8];
    bf16x8_to_f32(*reinterpret_cast<const uint4*>(p.x + static_cast<size_t>(pixel) * p.x_pitch + v * 8), f);
    const float* ca = p.ca + static_cast<size_t>(n) * p.c + v * 8;
    const float4 c0 = __ldg(reinterpret_cast<const float4*>(ca)), c1 = __ldg(reinterpret_cast<const float4*>(ca + 4));
    const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float t = f[i] * cc[i];
      s += t;
      m = fmaxf(m, t);
    }
  }
  for (int o = V >> 1; o > 0; o >>= 1) {     // the V lanes of a pixel are consecutive and V divides 32
    s += __shfl_xor_sync(0xffffffffu, s, o);
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  }
  if (on && v == 0) reinterpret_cast<float2*>(p.stats)[pixel] = make_float2(s / static_cast<float>(p.c), m);
}

__global__ void __launch_bounds__(kCbamThreads) cbam_spatial_kernel(const CbamParams p) {
  const long long idx = static_cast<long long>(blockIdx.x) * kCbamThreads + threadIdx.x;
  if (idx >= static_cast<long long>(p.batch) * p.hw) return;
  const int n = static_cast<int>(idx / p.hw);
  const int pix = static_cast<int>(idx - static_cast<long long>(n) * p.hw);
  const int y = pix / p.w, x = pix - y * p.w;
  const int k = p.ksize, pad = k >> 1;
  const float2* st = reinterpret_cast<const float2*>(p.stats) + static_cast<size_t>(n) * p.hw;
  float a = 0.0f;
  for (int dy = 0; dy < k; ++dy) {
    const int yy = y + dy - pad;
    if (yy < 0 || yy >= p.h) continue;
    for (int dx = 0; dx < k; ++dx) {
      const int xx = x + dx - pad;
      if (xx < 0 || xx >= p.w) continue;
      const float2 s = __ldg(st + yy * p.w + xx);
      a = fmaf(__ldg(p.conv + dy * k + dx), s.x, a);
      a = fmaf(__ldg(p.conv + k * k + dy * k + dx), s.y, a);
    }
  }
  p.sa[idx] = 1.0f / (1.0f + __expf(-a));
}

__global__ void __launch_bounds__(kCbamThreads) cbam_apply_vec_kernel(const CbamParams p, int V, int vshift) {
  const long long e = static_cast<long long>(blockIdx.x) * kCbamThreads + threadIdx.x;
  if (e >= static_cast<long long>(p.batch) * p.hw * V) return;
  const long long pixel = e >> vshift;
  const int v = static_cast<int>(e & (V - 1));
  const int n = static_cast<int>(pixel / p.hw);
  const float sa = __ldg(p.sa + pixel);
  float f[8];
  bf16x8_to_f32(*reinterpret_cast<const uint4*>(p.x + static_cast<size_t>(pixel) * p.x_pitch + v * 8), f);
  const float* ca = p.ca + static_cast<size_t>(n) * p.c + v * 8;
  const float4 c0 = __ldg(reinterpret_cast<const float4*>(ca)), c1 = __ldg(reinterpret_cast<const float4*>(ca + 4));
  const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
  uint4 o;
  o.x = pack_bf16x2(f[0] * cc[0] * sa, f[1] * cc[1] * sa);
  o.y = pack_bf16x2(f[2] * cc[2] * sa, f[3] * cc[3] * sa);
  o.z = pack_bf16x2(f[4] * cc[4] * sa, f[5] * cc[5] * sa);
  o.w = pack_bf16x2(f[6] * cc[6] * sa, f[7] * cc[7] * sa);
  *reinterpret_cast<uint4*>(p.out + static_cast<size_t>(pixel) * p.out_pitch + v * 8) = o;
}

static int cbam_chunks(int hw) { return (hw + kCbamChunkPx - 1) / kCbamChunkPx; }

}  // namespace tod

using namespace tod;

extern "C" int64_t tod_cbam_workspace_floats(int32_t batch, int32_t h, int32_t w, int32_t c) {
  if (batch <= 0 || h <= 0 || w <= 0 || c <= 0) return -1;
  const int64_t hw = static_cast<int64_t>(h) * w;
  return static_cast<int64_t>(batch) * (cbam_chunks(static_cast<int>(hw)) * 2 * c + c + hw * 3) + 64;
}

extern "C" int tod_cbam_nhwc_bf16(const tod_cbam_desc* d, void* stream) {
  TOD_CHECK_ARG(d != nullptr && d->d_x && d->d_out && d->d_fc1 && d->d_fc2 && d->d_conv && d->d_work, "cbam: null pointer");
  TOD_CHECK_ARG(d->batch > 0 && d->h > 0 && d->w > 0 && static_cast<long long>(d->batch) * d->h * d->w < (1ll << 31), "cbam: bad shape");
  TOD_CHECK_ARG(d->c >= 8 && d->c % 8 == 0 && d->c <= 2048 && d->hidden > 0 && d->hidden <= 256, "cbam: c %d hidden %d", d->c, d->hidden);
  TOD_CHECK_ARG(d->ksize >= 1 && d->ksize <= 15 && (d->ksize & 1), "cbam: kernel size %d", d->ksize);
  TOD_CHECK_ARG(d->x_pitch >= d->c && d->out_pitch >= d->c && d->x_pitch % 8 == 0 && d->out_pitch % 8 == 0, "cbam: pitches");
  TOD_CHECK_ARG((reinterpret_cast<uintptr_t>(d->d_x) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->d_out) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(d->d_work) & 15) == 0,
                "cbam: 16-byte alignment");
  CbamParams p;
  p.x = reinterpret_cast<const __nv_bfloat16*>(d->d_x);
  p.out = reinterpret_cast<__nv_bfloat16*>(d->d_out);
  p.fc1 = d->d_fc1;
  p.fc2 = d->d_fc2;
  p.conv = d->d_conv;
  p.batch = d->batch;
  p.h = d->h;
  p.w = d->w;
  p.hw = d->h * d->w;
  p.c = d->c;
  p.hidden = d->hidden;
  p.ksize = d->ksize;
  p.x_pitch = d->x_pitch;
  p.out_pitch = d->out_pitch;
  p.chunks = cbam_chunks(p.hw);
  const size_t n_partial = static_cast<size_t>(p.batch) * p.chunks * 2 * p.c;
  const size_t n_ca = (static_cast<size_t>(p.batch) * p.c + 3) / 4 * 4;
  p.partial = d->d_work;
  p.ca = d->d_work + (n_partial + 3) / 4 * 4;
  p.stats = p.ca + n_ca;
  p.sa = p.stats + static_cast<size_t>(p.batch) * p.hw * 2;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int V = p.c / 8;
  TOD_CHECK_ARG(V <= kCbamThreads, "cbam: too many channels for the pooling block");
  const int R = kCbamThreads / V;
  cbam_pool_kernel<<<dim3(p.chunks, p.batch), kCbamThreads, static_cast<size_t>(R) * V * 16 * sizeof(float), st>>>(p);
  TOD_CHECK_LAUNCH("cbam_pool_kernel launch");
  cbam_mlp_kernel<<<p.batch, kCbamThreads, (2 * p.c + 2 * p.hidden) * sizeof(float), st>>>(p);
  TOD_CHECK_LAUNCH("cbam_mlp_kernel launch");
  const unsigned blocks = static_cast<unsigned>((static_cast<long long>(p.batch) * p.hw + kCbamThreads - 1) / kCbamThreads);
  if (V <= 32 && (V & (V - 1)) == 0 && static_cast<long long>(p.batch) * p.hw * V < (1ll << 31) * kCbamThreads) {
    int vshift = 0;
    while ((1 << vshift) < V) ++vshift;
    const unsigned vblocks = static_cast<unsigned>((static_cast<long long>(p.batch) * p.hw * V + kCbamThreads - 1) / kCbamThreads);
    cbam_stats_vec_kernel<<<vblocks, kCbamThreads, 0, st>>>(p, V, vshift);
    TOD_CHECK_LAUNCH("cbam_stats_vec_kernel launch");
    cbam_spatial_kernel<<<blocks, kCbamThreads, 0, st>>>(p);
    TOD_CHECK_LAUNCH("cbam_spatial_kernel launch");
    cbam_apply_vec_kernel<<<vblocks, kCbamThreads, 0, st>>>(p, V, vshift);
    TOD_CHECK_LAUNCH("cbam_apply_vec_kernel launch");
    return TOD_OK;
  }
  cbam_stats_kernel<<<blocks, kCbamThreads, 0, st>>>(p);
  TOD_CHECK_LAUNCH("cbam_stats_kernel launch");
  cbam_apply_kernel<<<blocks, kCbamThreads, 0, st>>>(p);
  TOD_CHECK_LAUNCH("cbam_apply_kernel launch");
  return TOD_OK;
}

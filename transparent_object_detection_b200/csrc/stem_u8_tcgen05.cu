// Stem on the tensor cores from uint8 images: Conv2d(3, cout, 3, s2, p1) + folded BN + SiLU with the reference's
// `/255` pre-processing fused in.   uint8 NHWC (B, H, W, 3) in  ->  NHWC bf16 (B, H/2, W/2, out_pitch) out.
//
// Replaces backbone.stem (model/backbone.py:20; Conv.forward model/blocks.py:52-54) together with the host-side
// `np.array(image, float32) / 255.0` + HWC->CHW transpose in front of it (utils/callbacks.py:142-144,
// dataset/coco/get_map.py:52-60).  The CUDA-core stem (stem_conv.cu, 864 FMAs per output pixel) is FP32-FMA bound; here
// each 128-pixel tile builds its im2col rows (K = 27 padded to 32) in shared memory and two tcgen05 MMAs do the rest.
// Pixel values 0..255 are exact in bf16 and the 1/255 scale is folded into the bf16 weights, so this path adds no
// input quantisation on top of the weight rounding every other conv has.
//
// K index of a row = (kh * 3 + kw) * 3 + c  (9 contiguous bytes of an input row per kh); weights arrive as the folded
// f32 [cout][c * 9 + kh * 3 + kw] HOST array of tod_stem_conv_nchw_f32 and are re-ordered / scaled / rounded here.
#include <cstring>

#include "tma_host.cuh"

namespace tod {

constexpr int kStemU8Threads = 128;   // one accumulator row (= output pixel) per thread

template <int COUT>
struct StemU8Weights {
  float w[COUT * 27];   // [cout][c*9 + kh*3 + kw], BN folded
  float b[COUT];
};

template <int COUT>
__global__ void __launch_bounds__(kStemU8Threads) stem_u8_tcgen05_kernel(const uint8_t* __restrict__ x,
                                                                        __nv_bfloat16* __restrict__ out, int hin, int win,
                                                                        int out_pitch, long long total_pix,
                                                                        long long num_tiles,
                                                                        const __grid_constant__ StemU8Weights<COUT> wt) {
  constexpr uint32_t kCols = COUT <= 32 ? 32 : (COUT <= 64 ? 64 : 128);
  __shared__ __align__(1024) uint8_t a_tile[128 * 64];      // 128 rows x 32 bf16, SWIZZLE_64B
  __shared__ __align__(1024) uint8_t b_tile[COUT * 64];     // COUT rows x 32 bf16, SWIZZLE_64B
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ uint32_t tmem_base_smem;

  const int t = threadIdx.x, warp = t >> 5;
  const int hout = hin >> 1, wout = win >> 1;
  if (t == 0) {
    mbar_init(&mma_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_smem, kCols);
    tmem_relinquish();
  }

  // ---- B tile: weights / 255, bf16, row n at n*64 bytes, 16-byte chunk j stored at j ^ ((n >> 1) & 3)
  for (int i = t; i < COUT * 16; i += kStemU8Threads) {     // one u32 (two K values) per iteration
    const int n = i >> 4, kp = (i & 15) * 2;
    float v[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int k = kp + e;
      float wv = 0.f;
      if (k < 27) {
        const int tap = k / 3, c = k - tap * 3;
        wv = wt.w[n * 27 + c * 9 + tap] * (1.0f / 255.0f);
      }
      v[e] = wv;
    }
    const uint32_t chunk = (kp >> 3) ^ ((n >> 1) & 3);
    *reinterpret_cast<uint32_t*>(b_tile + n * 64 + chunk * 16 + (kp & 7) * 2) = pack_bf16x2(v[0], v[1]);
  }

  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base_smem;
  const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  const uint32_t hi = (512u >> 4) | (1u << 14) | (4u << 29);    // SBO = 8 rows x 64 B, SWIZZLE_64B
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(COUT >> 3) << 17) | ((128u >> 4) << 24);
  uint32_t phase = 0;

  // persistent loop: a handful of CTAs per SM interleave (im2col build | MMA | epilogue) of different tiles
#pragma unroll 1
  for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, phase ^= 1u) {
    // ---- A tile: this thread's im2col row
    const long long pix = tile * 128 + t;
    {
      uint32_t words[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) words[i] = 0;
      if (pix < total_pix) {
        const int ipix = static_cast<int>(pix);          // < 2^31, checked on the host
        const int r = ipix / wout;
        const int ow = ipix - r * wout;
        const int n = r / hout;
        const int oh = r - n * hout;
        const uint8_t* xn = x + static_cast<long long>(n) * hin * win * 3;
        float v[28];
        v[27] = 0.f;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const int ih = 2 * oh + kh - 1;
          const bool row_ok = ih >= 0 && ih < hin;
          const uint8_t* xr = xn + (static_cast<long long>(ih) * win + (2 * ow - 1)) * 3;
#pragma unroll
          for (int j = 0; j < 9; ++j) {
            const int iw = 2 * ow - 1 + j / 3;
            const bool ok = row_ok && iw >= 0 && iw < win;
            v[kh * 9 + j] = ok ? static_cast<float>(__ldg(xr + j)) : 0.f;
          }
        }
#pragma unroll
        for (int i = 0; i < 14; ++i) words[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t chunk = j ^ ((t >> 1) & 3);
        *reinterpret_cast<uint4*>(a_tile + t * 64 + chunk * 16) =
            make_uint4(words[4 * j], words[4 * j + 1], words[4 * j + 2], words[4 * j + 3]);
      }
    }
    fence_proxy_async_smem();      // generic-proxy smem writes -> visible to the tensor core (async proxy)
    tcgen05_fence_before();
    __syncthreads();               // also: every thread has drained the previous tile's accumulator
    tcgen05_fence_after();
    if (warp == 0) {
      if (elect_one()) {
        umma_bf16_k2(tmem, umma_desc_lo(smem_u32(a_tile)), hi, umma_desc_lo(smem_u32(b_tile)), hi, idesc, 0u);
        umma_commit(&mma_bar);
      }
      __syncwarp();
    }
    mbar_wait(&mma_bar, phase);    // MMAs complete: accumulator ready, a_tile free again
    tcgen05_fence_after();

    // ---- epilogue: thread t owns accumulator row t
    __nv_bfloat16* o = out + pix * out_pitch;
#pragma unroll
    for (int c0 = 0; c0 < COUT; c0 += 16) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(taddr + c0, v);
      tmem_ld_wait();
      if (pix < total_pix) {
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = silu_from_half(fmaf(__uint_as_float(v[j]), 0.5f, 0.5f * wt.b[c0 + j]));
#pragma unroll
        for (int j = 0; j < 16; j += 8) {
          uint4 ov;
          ov.x = pack_bf16x2(f[j], f[j + 1]);
          ov.y = pack_bf16x2(f[j + 2], f[j + 3]);
          ov.z = pack_bf16x2(f[j + 4], f[j + 5]);
          ov.w = pack_bf16x2(f[j + 6], f[j + 7]);
          *reinterpret_cast<uint4*>(o + c0 + j) = ov;
        }
      }
    }
    tcgen05_fence_before();        // the next iteration's __syncthreads orders these TMEM reads before its MMA
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc(tmem, kCols);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// TMA variant (the one the engine uses: win % 16 == 0, 16-byte aligned tensors).  A tile is 2 output rows x 64
// output columns; its 5 x 400-byte input patch arrives by ONE tensor-map load (out-of-bounds rows / words are
// zero-filled = the conv's padding) one tile ahead of its use, each thread builds its im2col row from shared memory with
// PRMT + FADD (no I2F), the epilogue is FFMA2 + one MUFU per element, and the finished tile leaves by one TMA store.
// No per-thread global address arithmetic is left (it was ~half of the issue slots of the kernel above).
constexpr int kStemTileW = 64;
constexpr int kPatchWords = 104;      // staged words per input row: bytes 6*ow0 - 16 .. 6*ow0 + 400 (measured: the innermost TMA
                                      // coordinate must start on a 16-byte boundary, else the load is an illegal instruction)
constexpr int kPatchRows = 5;         // input rows 2*oh0 - 1 .. 2*oh0 + 3

struct StemTma {
  CUtensorMap tm_in;    // uint32 [win*3/4, hin, batch, 1], box [104, 5, 1, 1]
  CUtensorMap tm_out;   // bf16 [cout, wout, hout, batch], box [stage channels, 64, 2, 1]
};

// uint8 -> float without the quarter-rate I2F: PRMT builds the bit pattern of 2^23 + b, one FADD removes the 2^23.
template <int BYTE>
__device__ __forceinline__ float byte_to_float(uint32_t word) {
  return __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7440 | BYTE)) - 8388608.0f;
}

template <int COUT>
__global__ void __launch_bounds__(kStemU8Threads) stem_u8_tma_kernel(const __grid_constant__ StemTma tm, int tiles_w, int tiles_h,
                                                                    int num_tiles,
                                                                    const __grid_constant__ StemU8Weights<COUT> wt, const TimelineTag tl) {
  const unsigned long long tl_t0 = (tl.buf != nullptr && threadIdx.x == 0) ? global_timer_ns() : 0ull;
  constexpr uint32_t kCols = COUT <= 32 ? 32 : (COUT <= 64 ? 64 : 128);
  constexpr int kStageC = COUT <= 64 ? COUT : (COUT % 64 == 0 ? 64 : 48);   // channels staged per epilogue pass
  constexpr int kRowBytes = kStageC * 2;                    // one staged output pixel
  // staging swizzle = the TMA store's swizzle mode for this row width (conflict-free 16-byte stores)
  constexpr int kSwz = kRowBytes == 128 ? 3 : (kRowBytes == 64 ? 2 : (kRowBytes == 32 ? 1 : 0));
  __shared__ __align__(1024) uint8_t a_tile[128 * 64];      // 128 rows x 32 bf16, SWIZZLE_64B
  __shared__ __align__(1024) uint8_t b_tile[COUT * 64];     // COUT rows x 32 bf16, SWIZZLE_64B
  __shared__ __align__(1024) uint8_t stage[128 * kRowBytes];
  __shared__ __align__(128) uint32_t patch[2][kPatchRows * kPatchWords + 24];   // 2176-byte slots (128-byte aligned)
  __shared__ __align__(16) float bias_h[COUT];              // bias / 2 (SiLU from the half argument)
  __shared__ __align__(8) uint64_t mma_bar, patch_full[2];
  __shared__ uint32_t tmem_base_smem;

  const int t = threadIdx.x, warp = t >> 5;
  const int tiles_per_img = tiles_w * tiles_h;
  auto issue_patch_load = [&](int tile, int buf) {          // one thread
    const int n = tile / tiles_per_img;
    const int rem = tile - n * tiles_per_img;
    const int ty = rem / tiles_w, tx = rem - ty * tiles_w;
    mbar_arrive_expect_tx(&patch_full[buf], kPatchRows * kPatchWords * 4);
    tma_load_4d(&tm.tm_in, &patch_full[buf], smem_u32(&patch[buf][0]), (3 * tx * kStemTileW) / 2 - 4, 4 * ty - 1, n, 0);
  };
  // warp 0: barriers + every TMA issue (one elected lane, warp-uniform branches); warp 1: TMEM owner + MMA issue
  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&tm.tm_in);
      tma_prefetch_desc(&tm.tm_out);
      mbar_init(&mma_bar, 1);
      mbar_init(&patch_full[0], 1);
      mbar_init(&patch_full[1], 1);
      fence_mbar_init();
    }
    __syncwarp();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, kCols);
    tmem_relinquish();
  }
  for (int i = t; i < COUT; i += kStemU8Threads) bias_h[i] = 0.5f * wt.b[i];

  // ---- B tile: weights / 255, bf16, row n at n*64 bytes, 16-byte chunk j stored at j ^ ((n >> 1) & 3)
  for (int i = t; i < COUT * 16; i += kStemU8Threads) {     // one u32 (two K values) per iteration
    const int n = i >> 4, kp = (i & 15) * 2;
    float v[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int k = kp + e;
      float wv = 0.f;
      if (k < 27) {
        const int tap = k / 3, c = k - tap * 3;
        wv = wt.w[n * 27 + c * 9 + tap] * (1.0f / 255.0f);
      }
      v[e] = wv;
    }
    const uint32_t chunk = (kp >> 3) ^ ((n >> 1) & 3);
    *reinterpret_cast<uint32_t*>(b_tile + n * 64 + chunk * 16 + (kp & 7) * 2) = pack_bf16x2(v[0], v[1]);
  }

  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 0 && static_cast<int>(blockIdx.x) < num_tiles) {   // first patch (barrier init is visible now)
    if (elect_one()) issue_patch_load(blockIdx.x, 0);
    __syncwarp();
  }
  const uint32_t tmem = tmem_base_smem;
  const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  const uint32_t hi = (512u >> 4) | (1u << 14) | (4u << 29);    // SBO = 8 rows x 64 B, SWIZZLE_64B
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(COUT >> 3) << 17) | ((128u >> 4) << 24);
  const int tr = t >> 6, tc = t & 63;                       // this thread's pixel inside the tile
  const int boff = 6 * tc + 13;                             // byte offset of input pixel 2*ow - 1 in a staged row
  const int pw0 = (2 * tr) * kPatchWords + (boff >> 2);     // first patch word of this thread's kh = 0 row
  const uint32_t sh = (boff & 3) * 8;                       // 8 or 24
  const uint32_t a_row = smem_u32(a_tile) + t * 64;
  const uint32_t a_swz = (t >> 1) & 3;
  const uint32_t st_row = smem_u32(stage) + t * kRowBytes;
  const uint32_t st_swz = kSwz == 3 ? (t & 7) : (kSwz == 2 ? ((t >> 1) & 3) : (kSwz == 1 ? ((t >> 2) & 1) : 0));
  const uint64_t half2 = pk2f(0.5f, 0.5f);

  // persistent loop: a handful of CTAs per SM interleave (im2col | MMA | epilogue | store) of different tiles
  int it = 0;
#pragma unroll 1
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
    const int buf = it & 1;
    if (warp == 0 && tile + static_cast<int>(gridDim.x) < num_tiles) {
      if (elect_one()) issue_patch_load(tile + gridDim.x, buf ^ 1);
      __syncwarp();
    }
    mbar_wait(&patch_full[buf], (it >> 1) & 1);
    // ---- A tile: this thread's im2col row (K = kh*9 + j, j = 9 contiguous input bytes) from the staged patch
    {
      uint32_t words[16];
      words[14] = 0;
      words[15] = 0;
      float v[28];
      v[27] = 0.f;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const uint32_t* pr = &patch[buf][pw0 + kh * kPatchWords];
        const uint32_t a = pr[0], b = pr[1], c = pr[2];
        const uint32_t v0 = __funnelshift_r(a, b, sh), v1 = __funnelshift_r(b, c, sh), v2 = c >> sh;
        v[kh * 9 + 0] = byte_to_float<0>(v0);
        v[kh * 9 + 1] = byte_to_float<1>(v0);
        v[kh * 9 + 2] = byte_to_float<2>(v0);
        v[kh * 9 + 3] = byte_to_float<3>(v0);
        v[kh * 9 + 4] = byte_to_float<0>(v1);
        v[kh * 9 + 5] = byte_to_float<1>(v1);
        v[kh * 9 + 6] = byte_to_float<2>(v1);
        v[kh * 9 + 7] = byte_to_float<3>(v1);
        v[kh * 9 + 8] = byte_to_float<0>(v2);
      }
#pragma unroll
      for (int i = 0; i < 14; ++i) words[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_row + ((j ^ a_swz) << 4)), "r"(words[4 * j]),
                     "r"(words[4 * j + 1]), "r"(words[4 * j + 2]), "r"(words[4 * j + 3])
                     : "memory");
    }
    if (t == 0) bulk_wait_read_all();   // the previous tile's TMA store (issued by this thread) has read the staging rows
    fence_proxy_async_smem();           // generic-proxy smem writes -> visible to the tensor core (async proxy)
    tcgen05_fence_before();
    __syncthreads();                    // also: every thread has drained the previous tile's accumulator
    tcgen05_fence_after();
    if (warp == 1) {
      if (elect_one()) {
        umma_bf16_k2(tmem, umma_desc_lo(smem_u32(a_tile)), hi, umma_desc_lo(smem_u32(b_tile)), hi, idesc, 0u);
        umma_commit(&mma_bar);
      }
      __syncwarp();
    }
    mbar_wait(&mma_bar, it & 1);        // MMAs complete: accumulator ready, a_tile free again
    tcgen05_fence_after();

    // ---- epilogue: thread t owns accumulator row t:  h = acc/2 + b/2,  silu = h + h * tanh(h)  (FFMA2 + one MUFU)
#pragma unroll
    for (int cb = 0; cb < COUT; cb += kStageC) {
      if (cb != 0) {                    // multi-pass (cout > 64): the previous pass's store must have read the staging rows
        if (t == 0) bulk_wait_read_all();
        __syncthreads();
      }
#pragma unroll
      for (int c0 = cb; c0 < cb + kStageC; c0 += 16) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(taddr + c0, v);
        tmem_ld_wait();
        uint32_t o[8];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 bb = *reinterpret_cast<const float4*>(bias_h + c0 + j);
          uint64_t f0 = ffma2_rn(pk2u32(v[j], v[j + 1]), half2, pk2f(bb.x, bb.y));
          uint64_t f1 = ffma2_rn(pk2u32(v[j + 2], v[j + 3]), half2, pk2f(bb.z, bb.w));
          float h0, h1, h2, h3, t0, t1, t2, t3;
          upk2f(f0, h0, h1);
          upk2f(f1, h2, h3);
          asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h0));
          asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h1));
          asm("tanh.approx.f32 %0, %1;" : "=f"(t2) : "f"(h2));
          asm("tanh.approx.f32 %0, %1;" : "=f"(t3) : "f"(h3));
          f0 = ffma2_rn(f0, pk2f(t0, t1), f0);
          f1 = ffma2_rn(f1, pk2f(t2, t3), f1);
          upk2f(f0, h0, h1);
          upk2f(f1, h2, h3);
          o[j >> 1] = pack_bf16x2(h0, h1);
          o[(j >> 1) + 1] = pack_bf16x2(h2, h3);
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const uint32_t chunk = (((c0 - cb) >> 3) + q) ^ st_swz;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_row + (chunk << 4)), "r"(o[4 * q]), "r"(o[4 * q + 1]),
                       "r"(o[4 * q + 2]), "r"(o[4 * q + 3])
                       : "memory");
        }
      }
      if (cb + kStageC >= COUT) tcgen05_fence_before();   // the next tile's __syncthreads orders these TMEM reads before its MMA
      fence_proxy_async_smem();
      __syncthreads();
      if (t == 0) {                     // one TMA store per tile (clipped at the image edges)
        const int n = tile / tiles_per_img;
        const int rem = tile - n * tiles_per_img;
        const int ty = rem / tiles_w, tx = rem - ty * tiles_w;
        tma_store_4d(&tm.tm_out, smem_u32(stage), cb, tx * kStemTileW, ty * 2, n);
        bulk_commit_group();
      }
      __syncwarp();                     // warp 0 reconverges before its next elect.sync / tcgen05.ld
    }
  }
  if (t == 0) bulk_wait_all();
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem, kCols);
  }
  if (t == 0) timeline_write(tl, tl_t0);
}

template <int COUT>
static int launch_stem_u8(const uint8_t* d_x, const float* h_w, const float* h_bias, void* d_out, int batch, int hin, int win,
                          int out_pitch, cudaStream_t st) {
  StemU8Weights<COUT> wt;
  memcpy(wt.w, h_w, sizeof(wt.w));
  if (h_bias) memcpy(wt.b, h_bias, sizeof(wt.b)); else memset(wt.b, 0, sizeof(wt.b));
  const long long total = static_cast<long long>(batch) * (hin / 2) * (win / 2);
  const long long tiles = (total + 127) / 128;
  TOD_CHECK_ARG(total < (1ll << 31), "stem_u8: too many output pixels");
  const int sms = num_sms();   // (the SM budget of the launch: tma_host.cuh)
  // CTAs per SM: bounded by TMEM (512 columns) and by what hides the build -> MMA -> epilogue latency chain
  constexpr int kCols = COUT <= 32 ? 32 : (COUT <= 64 ? 64 : 128);
  const int per_sm = 512 / kCols < 8 ? 512 / kCols : 8;
  const bool tma_ok = win % 16 == 0 && (reinterpret_cast<uintptr_t>(d_x) & 15) == 0 && (out_pitch * 2) % 16 == 0 &&
                      static_cast<long long>(batch) * ceil_div(win / 2, kStemTileW) * ceil_div(hin / 2, 2) < (1ll << 31);
  if (tma_ok) {
    constexpr int kStageC = COUT <= 64 ? COUT : (COUT % 64 == 0 ? 64 : 48);
    constexpr int kRowBytes = kStageC * 2;
    StemTma tm;
    int rc;
    {
      const uint64_t dims[4] = {static_cast<uint64_t>(win) * 3 / 4, static_cast<uint64_t>(hin), static_cast<uint64_t>(batch), 1};
      const uint64_t str[3] = {static_cast<uint64_t>(win) * 3, static_cast<uint64_t>(win) * 3 * hin,
                               static_cast<uint64_t>(win) * 3 * hin * batch};
      const uint32_t box[4] = {kPatchWords, kPatchRows, 1, 1};
      if ((rc = encode_map(&tm.tm_in, d_x, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_DATA_TYPE_UINT32)) != TOD_OK)
        return rc;
    }
    {
      const uint64_t px = static_cast<uint64_t>(out_pitch) * 2;
      const uint64_t dims[4] = {static_cast<uint64_t>(COUT), static_cast<uint64_t>(win / 2), static_cast<uint64_t>(hin / 2),
                                static_cast<uint64_t>(batch)};
      const uint64_t str[3] = {px, px * (win / 2), px * (win / 2) * (hin / 2)};
      const uint32_t box[4] = {kStageC, kStemTileW, 2, 1};
      const CUtensorMapSwizzle sw = kRowBytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                    : (kRowBytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                       : (kRowBytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE));
      if ((rc = encode_map(&tm.tm_out, d_out, 4, dims, str, box, sw)) != TOD_OK) return rc;
    }
    const int tiles_w = ceil_div(win / 2, kStemTileW), tiles_h = ceil_div(hin / 2, 2);
    const long long ntiles = static_cast<long long>(batch) * tiles_w * tiles_h;
    long long blocks = static_cast<long long>(sms) * per_sm;
    if (blocks > ntiles) blocks = ntiles;
    stem_u8_tma_kernel<COUT><<<static_cast<unsigned>(blocks), kStemU8Threads, 0, st>>>(tm, tiles_w, tiles_h,
                                                                                        static_cast<int>(ntiles), wt, timeline_tag("stem u8"));
    TOD_CHECK_LAUNCH("stem_u8_tma_kernel launch");
    return TOD_OK;
  }
  long long blocks = static_cast<long long>(sms) * per_sm;
  if (blocks > tiles) blocks = tiles;
  stem_u8_tcgen05_kernel<COUT><<<static_cast<unsigned>(blocks), kStemU8Threads, 0, st>>>(
      d_x, reinterpret_cast<__nv_bfloat16*>(d_out), hin, win, out_pitch, total, tiles, wt);
  TOD_CHECK_LAUNCH("stem_u8_tcgen05_kernel launch");
  return TOD_OK;
}

}  // namespace tod

using namespace tod;

extern "C" int tod_stem_conv_nhwc_u8(const uint8_t* d_x, const float* h_w, const float* h_bias, void* d_out, int32_t batch,
                                     int32_t hin, int32_t win, int32_t cout, int32_t out_pitch, void* stream) {
  TOD_CHECK_ARG(d_x && h_w && d_out, "stem_u8: null pointer");
  TOD_CHECK_ARG(batch > 0 && hin > 0 && win > 0 && hin % 2 == 0 && win % 2 == 0, "stem_u8: bad shape %d x %d x %d", batch,
                hin, win);
  TOD_CHECK_ARG(out_pitch >= cout && out_pitch % 8 == 0, "stem_u8: out_pitch %d", out_pitch);
  TOD_CHECK_ARG((reinterpret_cast<uintptr_t>(d_out) & 15) == 0, "stem_u8: output must be 16-byte aligned");
  auto st = static_cast<cudaStream_t>(stream);
  switch (cout) {
    case 16: return launch_stem_u8<16>(d_x, h_w, h_bias, d_out, batch, hin, win, out_pitch, st);
    case 32: return launch_stem_u8<32>(d_x, h_w, h_bias, d_out, batch, hin, win, out_pitch, st);
    case 48: return launch_stem_u8<48>(d_x, h_w, h_bias, d_out, batch, hin, win, out_pitch, st);
    case 64: return launch_stem_u8<64>(d_x, h_w, h_bias, d_out, batch, hin, win, out_pitch, st);
    case 96: return launch_stem_u8<96>(d_x, h_w, h_bias, d_out, batch, hin, win, out_pitch, st);
    case 128: return launch_stem_u8<128>(d_x, h_w, h_bias, d_out, batch, hin, win, out_pitch, st);
    default:
      set_error("stem_u8: unsupported cout %d (supported: 16, 32, 48, 64, 96, 128)", cout);
      return TOD_ERR_UNSUPPORTED;
  }
}

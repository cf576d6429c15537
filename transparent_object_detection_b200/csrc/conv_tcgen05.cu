// NHWC bf16 implicit-GEMM convolution on tcgen05 / TMEM fed by TMA (sm_100a).
//
// Replaces reference Conv.forward (model/blocks.py:52-54) with BN folded (fuse_conv, :160-187), the
// Bottleneck shortcut add (:80-82), the head's bare 1x1 convs (model/head.py:31,42) and -- by reading and
// writing channel-offset views -- every torch.cat/chunk on the path (see include/tod.h).
//
// GEMM view   D[M = pixels, N = cout] = A[M, K] * W[N, K]^T,   K = taps * cin_pad
//   A tile : 128 output pixels = a TH x TW patch of one image (3x3) or 128 consecutive pixels (1x1, "flat").
//            For tap (kh, kw) the A operand is the SAME patch shifted by (kh-1, kw-1): one TMA box load per
//            (tap, channel chunk); out-of-bounds rows/cols are zero-filled by TMA (= the conv's zero padding).
//            Stride 2 reads four parity sub-lattices of the input through four tensor maps.
//   W tile : [block_n, block_k] box of the packed [cout, K] weight matrix.
//   Both land in shared memory in the canonical K-major swizzled layout that tcgen05.mma consumes directly.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
//   warps 2..5 = epilogue (TMEM -> registers -> bias / upsample-add / SiLU / residual -> bf16|f32 -> global).
#include <cstring>
#include <mutex>

#include "tma_host.cuh"

namespace tod {

constexpr int kTileM = 128;
constexpr int kThreads = 320;  // TMA warp, MMA warp, 2 x 4 epilogue warps
constexpr int kMaxStages = 12;
constexpr int kSmemLimit = 226 * 1024;  // dynamic part; static barriers take <1 KB of the 227 KB opt-in limit

struct __align__(64) ConvKernelParams {
  CUtensorMap tm_a[4];
  CUtensorMap tm_w;
  // tiling
  int hout, wout;          // output spatial dims (flat mode: hout = 1, wout = total pixels)
  int th, tw;              // patch (th * tw <= 128)
  int tiles_w, tiles_h;    // tiles per image
  int n_tiles, total_tiles; // N tiles per M tile; M tiles x N tiles (N fastest)
  int cout, block_n, block_k;
  int num_taps, chunks_per_tap, num_stages;
  int tap_map[9], tap_dw[9], tap_dh[9];
  uint32_t stage_a_bytes, stage_b_bytes, tx_bytes;
  uint32_t idesc;          // tcgen05 instruction descriptor
  uint32_t desc_hi;        // upper 32 bits of the smem matrix descriptors (SBO, version, swizzle mode)
  uint32_t tmem_cols;
  // epilogue
  const float* bias;
  const __nv_bfloat16* residual;
  const float* upadd;
  void* out;
  int res_pitch, out_pitch;
  int up_h, up_w;          // real output dims for the upsample-add (flat mode needs them)
  int act, out_f32;
  int reverse;             // TOD_CONV_REVERSE: walk the tiles from the last to the first
  int nowait;              // experiment: skip the programmatic-launch wait (TOD_PDL_NOWAIT=1; results are wrong)
  TimelineTag tl;          // optional per-CTA start / end records (tools/timeline.py)
};

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t desc_hi) {
  // bits [0,14) start address >> 4, [16,30) leading byte offset >> 4 (unused for swizzled K-major: 1)
  return (static_cast<uint64_t>(desc_hi) << 32) | (1ull << 16) | ((smem_addr >> 4) & 0x3FFFu);
}

// One epilogue pass over 16 accumulator columns of one output row.
__device__ __forceinline__ void epilogue_store16(const ConvKernelParams& p, const uint32_t (&v)[16], int col,
                                                 long long pix, const float* up_ptr, const __nv_bfloat16* res_ptr) {
  float f[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
  if (p.bias) {
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col + j));
      f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
    }
  }
  if (up_ptr) {
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 u = __ldg(reinterpret_cast<const float4*>(up_ptr + col + j));
      f[j] += u.x; f[j + 1] += u.y; f[j + 2] += u.z; f[j + 3] += u.w;
    }
  }
  if (p.act == TOD_ACT_SILU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = silu_from_half(0.5f * f[j]);   // one MUFU.TANH instead of EX2 + RCP
  }
  if (res_ptr) {
#pragma unroll
    for (int j = 0; j < 16; j += 8) {
      const uint4 rv = __ldg(reinterpret_cast<const uint4*>(res_ptr + col + j));
      const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 rf = unpack_bf16x2(rw[t]);
        f[j + 2 * t] += rf.x;
        f[j + 2 * t + 1] += rf.y;
      }
    }
  }
  if (p.out_f32) {
    float* o = reinterpret_cast<float*>(p.out) + pix * p.out_pitch + col;
#pragma unroll
    for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
  } else {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.out_pitch + col;
#pragma unroll
    for (int j = 0; j < 16; j += 8) {
      uint4 ov;
      ov.x = pack_bf16x2(f[j], f[j + 1]);
      ov.y = pack_bf16x2(f[j + 2], f[j + 3]);
      ov.z = pack_bf16x2(f[j + 4], f[j + 5]);
      ov.w = pack_bf16x2(f[j + 6], f[j + 7]);
      *reinterpret_cast<uint4*>(o + j) = ov;
    }
  }
}

// Persistent kernel: every CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...  The smem ring and the two
// TMEM accumulator stages carry across tiles, so TMA loads of tile i+1/i+2, the MMAs of tile i+1 and the epilogue of
// tile i overlap.  Epilogue group g (4 warps) owns accumulator stage g = (local tile index) & 1.
__global__ void __launch_bounds__(kThreads, 1) conv_igemm_tcgen05(const __grid_constant__ ConvKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ unsigned long long tl_marks[2];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // 1024-byte aligned tile ring (swizzle atoms are 1024 B)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_bytes = p.stage_a_bytes + p.stage_b_bytes;
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const unsigned long long tl_t0 = (p.tl.buf != nullptr && threadIdx.x == 0) ? global_timer_ns() : 0ull;
  if (p.tl.buf != nullptr && threadIdx.x == 0) tl_marks[0] = tl_marks[1] = 0;
  const int num_chunks = p.num_taps * p.chunks_per_tap;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm_a[0]);
    tma_prefetch_desc(&p.tm_w);
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 4);   // one arrival per epilogue warp of the owning group
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, p.tmem_cols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  if (threadIdx.x == 0) pdl_launch_dependents();   // the next kernel's prologue may overlap this grid's tail

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one elected lane)
    if (elect_one()) {
      // every access of this grid to data of earlier kernels (activations, residual / upsample-add, the output buffer)
      // is ordered after this wait through the mbarrier chain that starts at the first load below
      if (!p.nowait) pdl_wait();
      if (p.tl.buf != nullptr) tl_marks[0] = global_timer_ns();
      int s = 0;         // ring slot and phase, continue across tiles
      uint32_t ph = 0;
      for (int tile_i = blockIdx.x; tile_i < p.total_tiles; tile_i += gridDim.x) {
        const int tile = p.reverse ? p.total_tiles - 1 - tile_i : tile_i;
        const int mt = tile / p.n_tiles;
        const int n0 = (tile - mt * p.n_tiles) * p.block_n;
        const int img = mt / tiles_per_img;
        const int trem = mt - img * tiles_per_img;
        const int tile_h = trem / p.tiles_w;
        const int h0 = tile_h * p.th;
        const int w0 = (trem - tile_h * p.tiles_w) * p.tw;
        int i = 0;
        for (int tap = 0; tap < p.num_taps; ++tap)
          for (int cc = 0; cc < p.chunks_per_tap; ++cc, ++i) {
            mbar_wait(&empty_bar[s], ph ^ 1u);
            mbar_arrive_expect_tx(&full_bar[s], p.tx_bytes);
            const uint32_t sa = smem_base + s * stage_bytes;
            tma_load_4d(&p.tm_a[p.tap_map[tap]], &full_bar[s], sa, cc * p.block_k, w0 + p.tap_dw[tap],
                        h0 + p.tap_dh[tap], img);
            tma_load_2d(&p.tm_w, &full_bar[s], sa + p.stage_a_bytes, i * p.block_k, n0);
            if (++s == p.num_stages) {
              s = 0;
              ph ^= 1u;
            }
          }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: one elected lane runs the whole role
    // (a warp-level wait + elect + __syncwarp per 4-MMA chunk drained the MMA queue between chunks)
    if (elect_one()) {
    const int ksteps = p.block_k >> 4;  // UMMA_K = 16 for bf16
    uint32_t lt = 0;                    // local tile index
    int s = 0;                          // ring slot and phase
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
      const uint32_t acc = lt & 1;
      mbar_wait(&tmem_empty_bar[acc], ((lt >> 1) & 1) ^ 1u);   // epilogue has drained this accumulator stage
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + acc * p.block_n;
      for (int i = 0; i < num_chunks; ++i) {
        mbar_wait(&full_bar[s], ph);
        tcgen05_fence_after();
        {
          const uint32_t sa = smem_base + s * stage_bytes;
          const uint32_t a_lo = umma_desc_lo(sa), b_lo = umma_desc_lo(sa + p.stage_a_bytes);
          // K steps of 16 bf16 = 32 bytes inside the swizzle atom: +2 in the (addr >> 4) field per step
          if (ksteps == 4)
            umma_bf16_k4(tmem_d, a_lo, p.desc_hi, b_lo, p.desc_hi, p.idesc, i != 0 ? 1u : 0u);
          else if (ksteps == 2)
            umma_bf16_k2(tmem_d, a_lo, p.desc_hi, b_lo, p.desc_hi, p.idesc, i != 0 ? 1u : 0u);
          else
            umma_bf16_k1(tmem_d, a_lo, p.desc_hi, b_lo, p.desc_hi, p.idesc, i != 0 ? 1u : 0u);
          umma_commit(&empty_bar[s]);                               // smem slot free once these MMAs retire
          if (i == num_chunks - 1) umma_commit(&tmem_full_bar[acc]);  // accumulator complete
        }
        if (++s == p.num_stages) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
    }   // elected lane
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue: 2 groups x 4 warps
    const int group = (warp - 2) >> 2;   // owns accumulator stage `group`
    const int q = warp & 3;              // a warp may only touch TMEM lanes [32*(warp%4), +32)
    const int r = q * 32 + lane;
    const int pth = r / p.tw;
    const int ptw = r - pth * p.tw;
    const bool in_patch = r < p.th * p.tw;
    uint32_t lt = 0;
    for (int tile_i = blockIdx.x; tile_i < p.total_tiles; tile_i += gridDim.x, ++lt) {
      if ((lt & 1) != static_cast<uint32_t>(group)) continue;
      const int tile = p.reverse ? p.total_tiles - 1 - tile_i : tile_i;
      const int mt = tile / p.n_tiles;
      const int n0 = (tile - mt * p.n_tiles) * p.block_n;
      const int img = mt / tiles_per_img;
      const int trem = mt - img * tiles_per_img;
      const int tile_h = trem / p.tiles_w;
      const int h = tile_h * p.th + pth;
      const int w = (trem - tile_h * p.tiles_w) * p.tw + ptw;
      const bool valid = in_patch && (h < p.hout) && (w < p.wout);
      const long long pix = (static_cast<long long>(img) * p.hout + h) * p.wout + w;
      const float* up_ptr = nullptr;
      if (p.upadd != nullptr && valid) {
        long long lp = pix;
        const int w_ = static_cast<int>(lp % p.up_w);
        lp /= p.up_w;
        const int h_ = static_cast<int>(lp % p.up_h);
        const long long n_ = lp / p.up_h;
        up_ptr = p.upadd + ((n_ * (p.up_h >> 1) + (h_ >> 1)) * (p.up_w >> 1) + (w_ >> 1)) * p.cout;
      }
      const __nv_bfloat16* res_ptr = p.residual ? p.residual + pix * p.res_pitch : nullptr;

      mbar_wait(&tmem_full_bar[group], (lt >> 1) & 1);
      if (p.tl.buf != nullptr && lt == 0 && r == 0) tl_marks[1] = global_timer_ns();
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + group * p.block_n;
      int c = 0;
      for (; c + 32 <= p.block_n; c += 32) {
        uint32_t v0[16], v1[16];
        tmem_ld_32x32b_x16(taddr + c, v0);
        tmem_ld_32x32b_x16(taddr + c + 16, v1);
        tmem_ld_wait();
        if (valid) {
          if (n0 + c < p.cout) epilogue_store16(p, v0, n0 + c, pix, up_ptr, res_ptr);
          if (n0 + c + 16 < p.cout) epilogue_store16(p, v1, n0 + c + 16, pix, up_ptr, res_ptr);
        }
      }
      if (c < p.block_n) {
        uint32_t v0[16];
        tmem_ld_32x32b_x16(taddr + c, v0);
        tmem_ld_wait();
        if (valid && n0 + c < p.cout) epilogue_store16(p, v0, n0 + c, pix, up_ptr, res_ptr);
      }
      // accumulator stage drained: hand it back to the MMA warp
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[group]);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
  if (threadIdx.x == 0) timeline_write(p.tl, tl_t0, tl_marks[0], tl_marks[1]);
}

// ------------------------------------------------------------------------------------------------
// CUDA-core evaluation of the same descriptor (tests only): one thread per (pixel, output channel).
struct SimtParams {
  const __nv_bfloat16* x;
  const __nv_bfloat16* w;
  const float* bias;
  const __nv_bfloat16* residual;
  const float* upadd;
  void* out;
  int batch, hin, win, cin, cout, hout, wout, ksize, stride, cin_pad;
  int x_pitch, res_pitch, out_pitch, act, out_f32;
};

__global__ void conv_simt_check(const SimtParams p) {
  const long long total = static_cast<long long>(p.batch) * p.hout * p.wout * p.cout;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int co = static_cast<int>(idx % p.cout);
  const long long pix = idx / p.cout;
  const int w = static_cast<int>(pix % p.wout);
  const int h = static_cast<int>((pix / p.wout) % p.hout);
  const int n = static_cast<int>(pix / (static_cast<long long>(p.wout) * p.hout));
  const int pad = p.ksize / 2;
  float acc = 0.f;
  for (int kh = 0; kh < p.ksize; ++kh)
    for (int kw = 0; kw < p.ksize; ++kw) {
      const int ih = h * p.stride + kh - pad, iw = w * p.stride + kw - pad;
      if (ih < 0 || ih >= p.hin || iw < 0 || iw >= p.win) continue;
      const __nv_bfloat16* xr = p.x + ((static_cast<long long>(n) * p.hin + ih) * p.win + iw) * p.x_pitch;
      const __nv_bfloat16* wr = p.w + (static_cast<long long>(co) * p.ksize * p.ksize + kh * p.ksize + kw) * p.cin_pad;
      for (int c = 0; c < p.cin; ++c) acc = fmaf(__bfloat162float(xr[c]), __bfloat162float(wr[c]), acc);
    }
  if (p.bias) acc += p.bias[co];
  if (p.upadd)
    acc += p.upadd[((static_cast<long long>(n) * (p.hout >> 1) + (h >> 1)) * (p.wout >> 1) + (w >> 1)) * p.cout + co];
  if (p.act == TOD_ACT_SILU) acc = acc / (1.0f + expf(-acc));
  if (p.residual) acc += __bfloat162float(p.residual[pix * p.res_pitch + co]);
  if (p.out_f32)
    reinterpret_cast<float*>(p.out)[pix * p.out_pitch + co] = acc;
  else
    reinterpret_cast<__nv_bfloat16*>(p.out)[pix * p.out_pitch + co] = __float2bfloat16_rn(acc);
}

// ------------------------------------------------------------------------------------------------ host
// patch shape minimising the number of 128-row tiles for an hout x wout map
static void pick_patch(int hout, int wout, int* th, int* tw) {
  long long best = -1;
  int bth = 1, btw = 1;
  const int wmax = wout < kTileM ? wout : kTileM;
  for (int t = 1; t <= wmax; ++t) {
    int hh = kTileM / t;
    if (hh > hout) hh = hout;
    if (hh > 256) hh = 256;
    const long long tiles = static_cast<long long>(ceil_div(wout, t)) * ceil_div(hout, hh);
    if (best < 0 || tiles < best || (tiles == best && t > btw)) {
      best = tiles;
      bth = hh;
      btw = t;
    }
  }
  *th = bth;
  *tw = btw;
}

static int validate(const tod_conv_desc* d, bool need_out = true) {
  TOD_CHECK_ARG(d != nullptr, "conv: null descriptor");
  TOD_CHECK_ARG(d->d_x && d->d_w && (d->d_out || !need_out), "conv: null x/w/out pointer");
  TOD_CHECK_ARG(d->batch > 0 && d->hin > 0 && d->win > 0, "conv: bad shape %d x %d x %d", d->batch, d->hin, d->win);
  TOD_CHECK_ARG((d->ksize == 1 && d->stride == 1) || (d->ksize == 3 && (d->stride == 1 || d->stride == 2)),
                "conv: unsupported ksize %d stride %d", d->ksize, d->stride);
  TOD_CHECK_ARG(d->cin >= 16 && d->cin % 16 == 0, "conv: cin %d must be a positive multiple of 16", d->cin);
  TOD_CHECK_ARG(d->cout >= 16 && d->cout % 16 == 0, "conv: cout %d must be a positive multiple of 16", d->cout);
  TOD_CHECK_ARG(d->x_pitch >= d->cin && d->x_pitch % 8 == 0, "conv: x_pitch %d", d->x_pitch);
  TOD_CHECK_ARG(d->out_pitch >= d->cout && d->out_pitch % 8 == 0, "conv: out_pitch %d", d->out_pitch);
  TOD_CHECK_ARG(!d->d_residual || (d->res_pitch >= d->cout && d->res_pitch % 8 == 0), "conv: res_pitch %d", d->res_pitch);
  TOD_CHECK_ARG(d->stride == 1 || (d->hin % 2 == 0 && d->win % 2 == 0), "conv: stride 2 needs even input dims");
  TOD_CHECK_ARG((reinterpret_cast<uintptr_t>(d->d_x) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->d_w) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(d->d_out) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(d->d_residual) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(d->d_upadd) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->d_bias) & 15) == 0,
                "conv: all device pointers must be 16-byte aligned");
  TOD_CHECK_ARG(d->act == TOD_ACT_NONE || d->act == TOD_ACT_SILU, "conv: bad act %d", d->act);
  TOD_CHECK_ARG(d->out_dtype == TOD_OUT_BF16 || d->out_dtype == TOD_OUT_F32, "conv: bad out_dtype %d", d->out_dtype);
  const int hout = d->hin / d->stride, wout = d->win / d->stride;
  TOD_CHECK_ARG(!d->d_upadd || (hout % 2 == 0 && wout % 2 == 0), "conv: upadd needs even output dims");
  return TOD_OK;
}

}  // namespace tod

using namespace tod;

extern "C" int tod_conv_weight_layout(int32_t cin, int32_t ksize, int32_t block_k_hint, int32_t* block_k,
                                      int32_t* cin_pad, int32_t* k_total) {
  TOD_CHECK_ARG(cin > 0 && (ksize == 1 || ksize == 3), "weight_layout: bad cin %d / ksize %d", cin, ksize);
  const int bk = pick_block_k(cin, block_k_hint);
  const int cp = round_up(cin, bk);
  if (block_k) *block_k = bk;
  if (cin_pad) *cin_pad = cp;
  if (k_total) *k_total = ksize * ksize * cp;
  return TOD_OK;
}

extern "C" int tod_conv2d_head_decode(const tod_conv_desc* d, const tod_head_fuse_desc* fuse, void* stream) {
  TOD_CHECK_ARG(fuse != nullptr, "fused head decode: null descriptor");
  int rc = validate(d, false);
  if (rc != TOD_OK) return rc;
  tod_conv_desc c = *d;
  c.out_dtype = TOD_OUT_F32;             // the logits stay f32 (TMEM) and never reach memory
  if (c.d_out == nullptr) c.d_out = const_cast<void*>(c.d_x);   // only used to build an (unused) output tensor map
  if (c.out_pitch < c.cout) c.out_pitch = c.cout;
  return conv_halo_launch(&c, stream, fuse);
}

extern "C" int tod_conv2d_tail1x1(const tod_conv_desc* d, const tod_conv_tail_desc* tail, void* stream) {
  TOD_CHECK_ARG(tail != nullptr, "conv tail: null descriptor");
  tod_conv_desc c = *d;
  if (c.d_out == nullptr) c.d_out = tail->d_out2;     // the intermediate is never stored
  if (c.out_pitch < c.cout) c.out_pitch = c.cout;
  int rc = validate(&c, false);
  if (rc != TOD_OK) return rc;
  return conv_halo_launch_tail(&c, tail, stream);
}

extern "C" int tod_conv2d_tail1x1_box_decode(const tod_conv_desc* d, const tod_conv_tail_desc* tail,
                                             const tod_head_fuse_desc* fuse, void* stream) {
  TOD_CHECK_ARG(tail != nullptr && fuse != nullptr, "conv tail + box decode: null descriptor");
  tod_conv_desc c = *d;
  if (c.d_out == nullptr) c.d_out = const_cast<void*>(c.d_x);   // neither the intermediate nor the logits are stored
  if (c.out_pitch < c.cout) c.out_pitch = c.cout;
  int rc = validate(&c, false);
  if (rc != TOD_OK) return rc;
  return conv_halo_launch_tail(&c, tail, stream, fuse);
}

extern "C" int tod_conv2d_nhwc_bf16(const tod_conv_desc* d, void* stream) {
  int rc = validate(d);
  if (rc != TOD_OK) return rc;
  // reserved[0]: kernel variant -- 0 auto, 1 per-tap loads (this file), 2 halo patches (conv_halo_tcgen05.cu).
  // Auto follows the per-layer measurements in profiles/r1_conv_bench_*.txt: the halo kernel (1 CTA/SM, patch reuse,
  // resident weights, specialised TMA-store epilogue) wins on maps >= 40 wide and on almost every 1x1 conv; the per-tap
  // kernel (2 CTAs/SM) still wins where 16x8 patches tile the map badly (20x20), on stride 2 with > 32 input channels
  // (four parity patches per chunk) and on the tiny 1x1 head outputs.
  int variant = d->reserved[0];
  if (variant == 0) {
    const int wout = d->win / d->stride;
    const long long mtot = static_cast<long long>(d->batch) * (d->hin / d->stride) * wout;
    bool halo;
    if (d->ksize == 3 && d->stride == 2) halo = d->cin <= 32;
    else if (d->ksize == 3) halo = wout >= 40 || flat_tiles_enabled();   // maps under 40 wide: row-flat tiles (halo kernel)
    else halo = !(mtot <= 32768 && d->cout <= 128);
    if (d->flags & TOD_CONV_PAIR_ON) halo = true;   // CTA pairs exist in the halo kernel only
    variant = halo ? 2 : 1;
  }
  if (variant == 2) return conv_halo_launch(d, stream);
  static PerDeviceOnce attr_once;   // the attribute is per device
  if (attr_once.needed()) {
    if ((rc = check_cuda(cudaFuncSetAttribute(conv_igemm_tcgen05, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit),
                         "cudaFuncSetAttribute(conv_igemm_tcgen05)")) != TOD_OK)
      return rc;
    attr_once.done();
  }

  ConvKernelParams p;
  memset(&p, 0, sizeof(p));
  const int hout = d->hin / d->stride, wout = d->win / d->stride;
  const int bk = pick_block_k(d->cin, d->block_k);
  const int cin_pad = round_up(d->cin, bk);
  const int taps = d->ksize * d->ksize;
  const int k_total = taps * cin_pad;
  const CUtensorMapSwizzle swz =
      bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (bk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  const uint32_t layout_type = bk == 64 ? 2u : (bk == 32 ? 4u : 6u);  // UMMA LayoutType: SW128 / SW64 / SW32
  const uint32_t row_bytes = bk * 2;
  const uint32_t sbo = 8 * row_bytes;  // 8-row core-matrix group stride

  // N tiling
  const int n_tiles = ceil_div(d->cout, 256);
  const int block_n = round_up(ceil_div(d->cout, n_tiles), 16);
  p.cout = d->cout;
  p.block_n = block_n;
  p.block_k = bk;
  p.num_taps = taps;
  p.chunks_per_tap = cin_pad / bk;

  const uint64_t px = static_cast<uint64_t>(d->x_pitch) * 2;  // bytes per input pixel
  const CUtensorMapL2promotion promo_in = l2_promotion_for(static_cast<uint64_t>(d->cin) * 2, px);
  if (d->ksize == 1) {
    // flat: 128 consecutive pixels per tile
    const long long mtot = static_cast<long long>(d->batch) * d->hin * d->win;
    p.hout = 1;
    p.wout = static_cast<int>(mtot);
    p.th = 1;
    p.tw = mtot < kTileM ? static_cast<int>(mtot) : kTileM;
    p.tiles_h = 1;
    p.tiles_w = ceil_div(static_cast<int>(mtot), p.tw);
    const uint64_t dims[4] = {static_cast<uint64_t>(d->cin), static_cast<uint64_t>(mtot), 1, 1};
    const uint64_t str[3] = {px, px * mtot, px * mtot};
    const uint32_t box[4] = {static_cast<uint32_t>(bk), static_cast<uint32_t>(p.tw), 1, 1};
    if ((rc = encode_map(&p.tm_a[0], d->d_x, 4, dims, str, box, swz, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, promo_in)) != TOD_OK) return rc;
    p.tap_map[0] = 0;
  } else {
    p.hout = hout;
    p.wout = wout;
    pick_patch(hout, wout, &p.th, &p.tw);
    p.tiles_h = ceil_div(hout, p.th);
    p.tiles_w = ceil_div(wout, p.tw);
    const uint32_t box[4] = {static_cast<uint32_t>(bk), static_cast<uint32_t>(p.tw), static_cast<uint32_t>(p.th), 1};
    if (d->stride == 1) {
      const uint64_t dims[4] = {static_cast<uint64_t>(d->cin), static_cast<uint64_t>(d->win),
                                static_cast<uint64_t>(d->hin), static_cast<uint64_t>(d->batch)};
      const uint64_t str[3] = {px, px * d->win, px * d->win * d->hin};
      if ((rc = encode_map(&p.tm_a[0], d->d_x, 4, dims, str, box, swz, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, promo_in)) != TOD_OK) return rc;
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw) {
          p.tap_map[kh * 3 + kw] = 0;
          p.tap_dh[kh * 3 + kw] = kh - 1;
          p.tap_dw[kh * 3 + kw] = kw - 1;
        }
    } else {
      // input row 2*oh + kh - 1:  kh=0 -> (oh-1, parity 1), kh=1 -> (oh, 0), kh=2 -> (oh, 1); same for columns
      const uint64_t dims[4] = {static_cast<uint64_t>(d->cin), static_cast<uint64_t>(d->win / 2),
                                static_cast<uint64_t>(d->hin / 2), static_cast<uint64_t>(d->batch)};
      const uint64_t str[3] = {px * 2, px * d->win * 2, px * d->win * d->hin};
      for (int ph = 0; ph < 2; ++ph)
        for (int pw = 0; pw < 2; ++pw) {
          const uint8_t* base = reinterpret_cast<const uint8_t*>(d->d_x) + (static_cast<uint64_t>(ph) * d->win + pw) * px;
          if ((rc = encode_map(&p.tm_a[ph * 2 + pw], base, 4, dims, str, box, swz, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, promo_in)) != TOD_OK) return rc;
        }
      const int par[3] = {1, 0, 1}, off[3] = {-1, 0, 0};
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw) {
          p.tap_map[kh * 3 + kw] = par[kh] * 2 + par[kw];
          p.tap_dh[kh * 3 + kw] = off[kh];
          p.tap_dw[kh * 3 + kw] = off[kw];
        }
    }
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(k_total), static_cast<uint64_t>(d->cout)};
    const uint64_t str[1] = {static_cast<uint64_t>(k_total) * 2};
    const uint32_t box[2] = {static_cast<uint32_t>(bk), static_cast<uint32_t>(block_n)};
    if ((rc = encode_map(&p.tm_w, d->d_w, 2, dims, str, box, swz)) != TOD_OK) return rc;
  }

  p.stage_a_bytes = kTileM * row_bytes;
  p.stage_b_bytes = round_up(block_n * row_bytes, 1024);
  p.tx_bytes = static_cast<uint32_t>(p.th * p.tw) * row_bytes + block_n * row_bytes;
  const int num_chunks = taps * p.chunks_per_tap;
  const uint32_t stage_bytes = p.stage_a_bytes + p.stage_b_bytes;
  // two CTAs per SM when both fit (TMEM: 2 CTAs x 2 stages x block_n columns <= 512); else one CTA with a deep ring
  const int ctas_per_sm = block_n <= 128 ? 2 : 1;
  int stages = d->num_stages;
  if (stages <= 0) {
    const uint32_t budget = ctas_per_sm == 2 ? 108 * 1024 : 220 * 1024;
    stages = budget / stage_bytes;
  }
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 1) stages = 1;
  while (stages > 1 && stages * stage_bytes + 1024 > (uint32_t)kSmemLimit) --stages;
  p.num_stages = stages;
  const size_t smem = static_cast<size_t>(stages) * stage_bytes + 1024;

  // instruction descriptor: D=f32, A=B=bf16, both K-major, N = block_n, M = 128
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(block_n >> 3) << 17) |
            (static_cast<uint32_t>(kTileM >> 4) << 24);
  p.desc_hi = (sbo >> 4) | (1u << 14) | (layout_type << 29);
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(2 * block_n)) cols <<= 1;   // two accumulator stages
  p.tmem_cols = cols;

  p.bias = d->d_bias;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(d->d_residual);
  p.upadd = d->d_upadd;
  p.out = d->d_out;
  p.res_pitch = d->res_pitch;
  p.out_pitch = d->out_pitch;
  p.up_h = hout;
  p.up_w = wout;
  p.act = d->act;
  p.out_f32 = d->out_dtype == TOD_OUT_F32;
  p.reverse = (d->flags & TOD_CONV_REVERSE) ? 1 : 0;
  p.nowait = pdl_nowait();
  {
    char nm[48];
    snprintf(nm, sizeof(nm), "tapconv %d>%d k%d s%d @%dx%d", d->cin, d->cout, d->ksize, d->stride, hout, wout);
    p.tl = timeline_tag(nm);
  }

  const long long m_tiles = static_cast<long long>(d->ksize == 1 ? 1 : d->batch) * p.tiles_h * p.tiles_w;
  p.n_tiles = ceil_div(d->cout, block_n);
  TOD_CHECK_ARG(m_tiles > 0 && m_tiles * p.n_tiles < (1ll << 31), "conv: too many tiles");
  p.total_tiles = static_cast<int>(m_tiles * p.n_tiles);
  const int slots = num_sms() * ctas_per_sm;
  const int grid = p.total_tiles < slots ? p.total_tiles : slots;
  if ((rc = check_cuda(launch_pdl(conv_igemm_tcgen05, dim3(grid), dim3(kThreads), smem, static_cast<cudaStream_t>(stream), p),
                       "conv_igemm_tcgen05 launch")) != TOD_OK)
    return rc;
  TOD_CHECK_LAUNCH("conv_igemm_tcgen05 launch");
  return TOD_OK;
}

extern "C" int tod_conv2d_nhwc_bf16_simt_check(const tod_conv_desc* d, void* stream) {
  int rc = validate(d);
  if (rc != TOD_OK) return rc;
  SimtParams p;
  p.x = reinterpret_cast<const __nv_bfloat16*>(d->d_x);
  p.w = reinterpret_cast<const __nv_bfloat16*>(d->d_w);
  p.bias = d->d_bias;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(d->d_residual);
  p.upadd = d->d_upadd;
  p.out = d->d_out;
  p.batch = d->batch; p.hin = d->hin; p.win = d->win; p.cin = d->cin; p.cout = d->cout;
  p.hout = d->hin / d->stride; p.wout = d->win / d->stride;
  p.ksize = d->ksize; p.stride = d->stride;
  p.cin_pad = round_up(d->cin, pick_block_k(d->cin, d->block_k));
  p.x_pitch = d->x_pitch; p.res_pitch = d->res_pitch; p.out_pitch = d->out_pitch;
  p.act = d->act; p.out_f32 = d->out_dtype == TOD_OUT_F32;
  const long long total = static_cast<long long>(p.batch) * p.hout * p.wout * p.cout;
  const int threads = 256;
  conv_simt_check<<<static_cast<unsigned>((total + threads - 1) / threads), threads, 0,
                    static_cast<cudaStream_t>(stream)>>>(p);
  TOD_CHECK_LAUNCH("conv_simt_check launch");
  return TOD_OK;
}

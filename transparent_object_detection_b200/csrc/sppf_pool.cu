// SPPF pooling: three chained MaxPool2d(k=5, s=1, p=2) over an NHWC bf16 plane held in shared memory.
// Replaces SPPF.forward's `y.extend(self.m(y[-1]) for _ in range(3))` + torch.cat (model/blocks.py:139-141):
// the input already sits in channels [0,c) of the 4c-wide concat buffer (written there by cv1) and the three
// pooled planes are written to channels [c,2c), [2c,3c), [3c,4c) -- no cat, no intermediate round trip.
// HBM-bound: reads c*H*W*2 bytes, writes 3x that.  One CTA = one image x one group of CG channels; each 5x5
// max is done separably (row pass, column pass) with 8-channel (16-byte) vectors; -inf padding == clipped windows.
#include <cstdlib>

#include "tma_host.cuh"

namespace tod {

constexpr int kPoolThreads = 256;

// TOD_POOL_TMA=0: the LDG / STG kernels below instead of the TMA kernel (A/B measurements; both are exact)
static bool pool_tma_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("TOD_POOL_TMA");
    on = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

__device__ __forceinline__ uint4 max_bf16x8(uint4 a, uint4 b) {
  uint4 r;
  __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
  for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
  return r;
}

// VEC = number of 8-channel vectors per pixel handled by one CTA (CG = 8 * VEC channels)
template <int VEC>
__global__ void __launch_bounds__(kPoolThreads) sppf_pool_kernel(__nv_bfloat16* __restrict__ buf, int h, int w, int c,
                                                                 int pitch) {
  extern __shared__ uint4 pool_smem[];
  uint4* b0 = pool_smem;
  uint4* b1 = pool_smem + static_cast<size_t>(h) * w * VEC;
  const int groups = c / (8 * VEC);
  const int n = blockIdx.x / groups;
  const int c0 = (blockIdx.x - n * groups) * 8 * VEC;
  const int total = h * w * VEC;
  __nv_bfloat16* img = buf + static_cast<size_t>(n) * h * w * pitch + c0;

  for (int i = threadIdx.x; i < total; i += kPoolThreads) {
    const int px = i / VEC, v = i - px * VEC;
    b0[i] = *reinterpret_cast<const uint4*>(img + static_cast<size_t>(px) * pitch + v * 8);
  }
  __syncthreads();
  for (int stage = 1; stage <= 3; ++stage) {
    // row pass b0 -> b1
    for (int i = threadIdx.x; i < total; i += kPoolThreads) {
      const int px = i / VEC, v = i - px * VEC;
      const int y = px / w, x = px - y * w;
      const int xa = x - 2 < 0 ? 0 : x - 2, xb = x + 2 >= w ? w - 1 : x + 2;
      uint4 m = b0[(y * w + xa) * VEC + v];
      for (int xx = xa + 1; xx <= xb; ++xx) m = max_bf16x8(m, b0[(y * w + xx) * VEC + v]);
      b1[i] = m;
    }
    __syncthreads();
    // column pass b1 -> b0 and out to channel slot `stage`
    for (int i = threadIdx.x; i < total; i += kPoolThreads) {
      const int px = i / VEC, v = i - px * VEC;
      const int y = px / w, x = px - y * w;
      const int ya = y - 2 < 0 ? 0 : y - 2, yb = y + 2 >= h ? h - 1 : y + 2;
      uint4 m = b1[(ya * w + x) * VEC + v];
      for (int yy = ya + 1; yy <= yb; ++yy) m = max_bf16x8(m, b1[(yy * w + x) * VEC + v]);
      b0[i] = m;
      *reinterpret_cast<uint4*>(img + static_cast<size_t>(px) * pitch + stage * c + v * 8) = m;
    }
    __syncthreads();
  }
}

// Planes of up to kFastElems * 256 (pixel, vector) elements (every plane of the 640^2 detector): each thread owns
// the same <= 4 elements in every pass, and their clamped neighbour indices (a repeated edge element is harmless under
// max == the reference's implicit -inf padding) are computed ONCE, so an element-pass is 5 shared-memory loads and 16
// packed bf16 max.  The generic kernel above re-derives (y, x) with two integer divisions per element per pass and ran
// at ~95 instructions per element-pass: instruction-bound (ncu: issue-active 53 %), 32 us for 52 MB.
constexpr int kFastElems = 4;

template <int VEC>
__global__ void __launch_bounds__(kPoolThreads) sppf_pool_fast_kernel(__nv_bfloat16* __restrict__ buf, int h, int w, int c,
                                                                      int pitch, const TimelineTag tl) {
  const unsigned long long tl_t0 = (tl.buf != nullptr && threadIdx.x == 0) ? global_timer_ns() : 0ull;
  extern __shared__ uint4 pool_smem[];
  const int total = h * w * VEC;            // <= kFastElems * kPoolThreads (host)
  uint4* b0 = pool_smem;
  uint4* b1 = pool_smem + total;
  const int groups = c / (8 * VEC);
  const int n = blockIdx.x / groups;
  const int c0 = (blockIdx.x - n * groups) * 8 * VEC;
  __nv_bfloat16* img = buf + static_cast<size_t>(n) * h * w * pitch + c0;

  int self[kFastElems], nb_row[kFastElems][4], nb_col[kFastElems][4], goff[kFastElems];
  bool on[kFastElems];
#pragma unroll
  for (int e = 0; e < kFastElems; ++e) {
    const int i = e * kPoolThreads + threadIdx.x;
    on[e] = i < total;
    self[e] = on[e] ? i : 0;
    const int px = self[e] / VEC, v = self[e] - px * VEC;
    const int y = px / w, x = px - y * w;
    goff[e] = px * pitch + v * 8;
    const int dd[4] = {-2, -1, 1, 2};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      nb_row[e][k] = (y * w + min(max(x + dd[k], 0), w - 1)) * VEC + v;
      nb_col[e][k] = (min(max(y + dd[k], 0), h - 1) * w + x) * VEC + v;
    }
  }
#pragma unroll
  for (int e = 0; e < kFastElems; ++e)
    if (on[e]) b0[self[e]] = *reinterpret_cast<const uint4*>(img + goff[e]);
  __syncthreads();
#pragma unroll 1
  for (int stage = 1; stage <= 3; ++stage) {
#pragma unroll
    for (int e = 0; e < kFastElems; ++e) {            // row pass b0 -> b1
      if (on[e]) {
        uint4 m = b0[self[e]];
#pragma unroll
        for (int k = 0; k < 4; ++k) m = max_bf16x8(m, b0[nb_row[e][k]]);
        b1[self[e]] = m;
      }
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < kFastElems; ++e) {            // column pass b1 -> b0 and out to channel slot `stage`
      if (on[e]) {
        uint4 m = b1[self[e]];
#pragma unroll
        for (int k = 0; k < 4; ++k) m = max_bf16x8(m, b1[nb_col[e][k]]);
        b0[self[e]] = m;
        *reinterpret_cast<uint4*>(img + goff[e] + stage * c) = m;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) timeline_write(tl, tl_t0);
}

// TMA version (every plane of the detector at 640^2 and 1280^2).  The kernels above turned out to be bound by shared-memory
// bandwidth, not HBM: a separable 5-tap max that re-reads its five inputs per output moves 96 bytes of shared memory per
// 16-byte element and pass (ncu: 27 us for 52 MB = 28 % of the HBM peak).  Here one thread owns a whole row (then a whole
// column) of one 8-channel vector and keeps the window in registers -- 1.4 shared-memory loads and one store per element
// and pass, three packed maxima via m2[i] = max(x[i], x[i+1]), m5[i] = max(m2[i], m2[i+2], x[i+4]) -- the plane arrives
// by ONE tensor-map load and every pooled plane leaves by one TMA store straight from the
// buffer the next stage reads.  One CTA = one image x one group of 8 * VEC channels.
struct PoolTma {
  CUtensorMap tm;   // [batch, h, w, pitch] bf16, box {8 * VEC channels, w, h, 1}
};

template <int VEC>
__global__ void __launch_bounds__(kPoolThreads) sppf_pool_tma_kernel(const __grid_constant__ PoolTma tm, int h, int w, int c,
                                                                     const TimelineTag tl) {
  const unsigned long long tl_t0 = (tl.buf != nullptr && threadIdx.x == 0) ? global_timer_ns() : 0ull;
  extern __shared__ uint8_t pool_raw[];
  __shared__ __align__(8) uint64_t full_bar;
  const uint32_t base = (smem_u32(pool_raw) + 127u) & ~127u;
  const int total = h * w * VEC;                       // uint4 elements of one plane
  const uint32_t b0 = base, b1 = base + static_cast<uint32_t>(total) * 16u;
  const int groups = c / (8 * VEC);
  const int n = blockIdx.x / groups;
  const int c0 = (blockIdx.x - n * groups) * 8 * VEC;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm.tm);
    mbar_init(&full_bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&full_bar, static_cast<uint32_t>(total) * 16u);
    tma_load_4d(&tm.tm, &full_bar, b0, c0, 0, 0, n);
  }
  mbar_wait(&full_bar, 0);
  auto lds = [](uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
  };
  auto sts = [](uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  };
  // one pass along a line of `len` elements `stride` bytes apart: out[i] = max(in[i-2 .. i+2]) clipped to the line (a
  // clipped window == the reference's implicit -inf padding, model/blocks.py:135; a repeated edge element is harmless
  // under max).  A task is ten consecutive outputs of one line, from fourteen loads issued together: the loads are
  // independent, so the shared-memory latency is paid once per task instead of once per element, and a 20-element line is
  // two tasks (every thread of the CTA has work in every pass).
  constexpr int CH = 10;
  auto chunk_pass = [&](uint32_t src, uint32_t dst, int len, uint32_t stride, int base_i) {
    uint4 x[CH + 4];
#pragma unroll
    for (int j = 0; j < CH + 4; ++j) {
      const int idx = min(max(base_i + j - 2, 0), len - 1);
      x[j] = lds(src + static_cast<uint32_t>(idx) * stride);
    }
    uint4 m2[CH + 3];
#pragma unroll
    for (int j = 0; j < CH + 3; ++j) m2[j] = max_bf16x8(x[j], x[j + 1]);
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      if (base_i + j < len) {
        const uint4 m = max_bf16x8(max_bf16x8(m2[j], m2[j + 2]), x[j + 4]);
        sts(dst + static_cast<uint32_t>(base_i + j) * stride, m);
      }
    }
  };
  const uint32_t px_bytes = VEC * 16u, row_bytes = static_cast<uint32_t>(w) * px_bytes;
#pragma unroll 1
  for (int stage = 1; stage <= 3; ++stage) {
    // row pass b0 -> b1: task = (row y, vector v)
    const int row_chunks = (w + CH - 1) / CH, col_chunks = (h + CH - 1) / CH;
    for (int tsk = threadIdx.x; tsk < h * VEC * row_chunks; tsk += kPoolThreads) {
      const int line = tsk % (h * VEC), chunk = tsk / (h * VEC);
      const int y = line / VEC, v = line - y * VEC;
      const uint32_t off = static_cast<uint32_t>(y) * row_bytes + v * 16u;
      chunk_pass(b0 + off, b1 + off, w, px_bytes, chunk * CH);
    }
    if (threadIdx.x == 0) bulk_wait_read_all();   // the previous stage's store has finished reading b0
    __syncthreads();
    // column pass b1 -> b0: task = (column x, vector v)
    for (int tsk = threadIdx.x; tsk < w * VEC * col_chunks; tsk += kPoolThreads) {
      const int line = tsk % (w * VEC), chunk = tsk / (w * VEC);
      const uint32_t off = static_cast<uint32_t>(line) * 16u;   // (x * VEC + v) * 16
      chunk_pass(b1 + off, b0 + off, h, row_bytes, chunk * CH);
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {   // pooled plane `stage` -> channel slot `stage` of the concat buffer
      tma_store_4d(&tm.tm, b0, stage * c + c0, 0, 0, n);
      bulk_commit_group();
    }
  }
  if (threadIdx.x == 0) {
    bulk_wait_read_all();
    timeline_write(tl, tl_t0);
  }
}

}  // namespace tod

using namespace tod;

extern "C" int tod_sppf_pool_nhwc_bf16(void* d_buf, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t pitch,
                                       void* stream) {
  TOD_CHECK_ARG(d_buf != nullptr, "sppf_pool: null buffer");
  TOD_CHECK_ARG(batch > 0 && h > 0 && w > 0, "sppf_pool: bad shape");
  TOD_CHECK_ARG(c > 0 && c % 8 == 0 && pitch >= 4 * c && pitch % 8 == 0, "sppf_pool: c %d pitch %d", c, pitch);
  TOD_CHECK_ARG((reinterpret_cast<uintptr_t>(d_buf) & 15) == 0, "sppf_pool: buffer must be 16-byte aligned");
  static PerDeviceOnce attr_once;   // the attribute is per device
  if (attr_once.needed()) {
    int rc = check_cuda(cudaFuncSetAttribute(sppf_pool_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024),
                        "cudaFuncSetAttribute(sppf_pool<2>)");
    if (rc != TOD_OK) return rc;
    rc = check_cuda(cudaFuncSetAttribute(sppf_pool_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024),
                    "cudaFuncSetAttribute(sppf_pool<1>)");
    if (rc != TOD_OK) return rc;
    attr_once.done();
  }
  const size_t plane16 = static_cast<size_t>(h) * w * 16;  // bytes for one 8-channel vector plane
  auto st = static_cast<cudaStream_t>(stream);
  // TMA path: the widest channel group whose two plane buffers stay under ~100 KB (two CTAs per SM)
  if (pool_tma_enabled() && h <= 256 && w <= 256 && h >= 3 && w >= 3) {
    int vec = 0;
    static int vec_cap = -1;   // TOD_POOL_VEC: cap on the channel-group width (tools)
    if (vec_cap < 0) {
      const char* e = getenv("TOD_POOL_VEC");
      vec_cap = e != nullptr ? atoi(e) : 8;
    }
    for (int v : {8, 4, 2})
      if (v <= vec_cap && c % (8 * v) == 0 && 2 * plane16 * v + 256 <= 104 * 1024) {
        vec = v;
        break;
      }
    if (vec != 0) {
      PoolTma tm;
      const uint64_t px = static_cast<uint64_t>(pitch) * 2;
      const uint64_t dims[4] = {static_cast<uint64_t>(pitch), static_cast<uint64_t>(w), static_cast<uint64_t>(h),
                                static_cast<uint64_t>(batch)};
      const uint64_t str[3] = {px, px * w, px * w * h};
      const uint32_t box[4] = {static_cast<uint32_t>(8 * vec), static_cast<uint32_t>(w), static_cast<uint32_t>(h), 1};
      int rc = encode_map(&tm.tm, d_buf, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                          CU_TENSOR_MAP_L2_PROMOTION_NONE);
      if (rc != TOD_OK) return rc;
      const size_t smem = 2 * plane16 * vec + 256;
      const unsigned grid = static_cast<unsigned>(batch) * (c / (8 * vec));
      static PerDeviceOnce tma_attr_once;
      if (tma_attr_once.needed()) {
        rc = check_cuda(cudaFuncSetAttribute(sppf_pool_tma_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 104 * 1024), "attr pool<8>");
        if (rc == TOD_OK) rc = check_cuda(cudaFuncSetAttribute(sppf_pool_tma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 104 * 1024), "attr pool<4>");
        if (rc == TOD_OK) rc = check_cuda(cudaFuncSetAttribute(sppf_pool_tma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 104 * 1024), "attr pool<2>");
        if (rc != TOD_OK) return rc;
        tma_attr_once.done();
      }
      const TimelineTag tag = timeline_tag("sppf pool");
      if (vec == 8) sppf_pool_tma_kernel<8><<<grid, kPoolThreads, smem, st>>>(tm, h, w, c, tag);
      else if (vec == 4) sppf_pool_tma_kernel<4><<<grid, kPoolThreads, smem, st>>>(tm, h, w, c, tag);
      else sppf_pool_tma_kernel<2><<<grid, kPoolThreads, smem, st>>>(tm, h, w, c, tag);
      TOD_CHECK_LAUNCH("sppf_pool_tma_kernel launch");
      return TOD_OK;
    }
  }
  if (c % 16 == 0 && h * w * 2 <= kFastElems * kPoolThreads) {
    sppf_pool_fast_kernel<2><<<batch * (c / 16), kPoolThreads, 2 * 2 * plane16, st>>>(
        reinterpret_cast<__nv_bfloat16*>(d_buf), h, w, c, pitch, timeline_tag("sppf pool"));
    TOD_CHECK_LAUNCH("sppf_pool_fast_kernel launch");
    return TOD_OK;
  }
  if (h * w <= kFastElems * kPoolThreads) {
    sppf_pool_fast_kernel<1><<<batch * (c / 8), kPoolThreads, 2 * plane16, st>>>(reinterpret_cast<__nv_bfloat16*>(d_buf), h, w,
                                                                                c, pitch, timeline_tag("sppf pool"));
    TOD_CHECK_LAUNCH("sppf_pool_fast_kernel launch");
    return TOD_OK;
  }
  if (c % 16 == 0 && 2 * 2 * plane16 <= 200 * 1024) {
    sppf_pool_kernel<2><<<batch * (c / 16), kPoolThreads, 2 * 2 * plane16, st>>>(
        reinterpret_cast<__nv_bfloat16*>(d_buf), h, w, c, pitch);
  } else {
    TOD_CHECK_ARG(2 * plane16 <= 200 * 1024, "sppf_pool: %d x %d plane does not fit shared memory", h, w);
    sppf_pool_kernel<1><<<batch * (c / 8), kPoolThreads, 2 * plane16, st>>>(reinterpret_cast<__nv_bfloat16*>(d_buf), h,
                                                                            w, c, pitch);
  }
  TOD_CHECK_LAUNCH("sppf_pool_kernel launch");
  return TOD_OK;
}

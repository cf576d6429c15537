// Class-wise greedy NMS, bit-exact with the reference's per-image / per-class torchvision.ops.nms loops.
// Replaces DecodeBox.non_max_suppression (utils/bbox_utils.py:144-175) and torchvision.ops.nms (:172).
//
//   nms_prepare_dense   (dense reference tensor only) xywh -> corners in place + class max        [:144-153]
//   nms_sort_kernel     conf filter (>=, float32), 64-bit key = (class asc | score desc | anchor asc),
//                       bitonic sort per image in shared memory, class segments found by binary search
//                       -> the reference's output ORDER (unique() ascending :164, nms score-descending)   [:156-172]
//   nms_segment_kernel  one warp per (image, class) segment: 32-box tiles; inside a tile every lane builds its
//                       32-bit IoU bitmask and the greedy scan is resolved with warp shuffles; kept boxes of the
//                       tile then suppress the rest of the segment.  IoU arithmetic = torchvision's, in float32
//                       with IEEE division and no FMA contraction (file compiled with -fmad=false).
//   nms_compact_kernel  ordered compaction of the survivors -> keep indices, counts, [x1,y1,x2,y2,conf,cls] rows  [:173-175]
// The work is latency/compare-bound, not HBM-bound: ~24 B per candidate in, 8..28 B per kept box out.
#include <cmath>
#include <cstring>

#include "tod_common.cuh"

namespace tod {

constexpr int kSortThreads = 1024;
constexpr int kSortSmemKeys = 16384;  // 128 KB of 64-bit keys
constexpr int kSegWarps = 16;
constexpr int kSegCtasPerImage = 16;
constexpr int kIdxBits = 20, kScoreBits = 32;  // key = cls[12] | ~score[32] | idx[20]
constexpr unsigned long long kIdxMask = (1ull << kIdxBits) - 1;

struct NmsWork {
  int* order;              // [B][A] anchor index, sorted
  int* seg_start;          // [B][A]
  int* seg_len;            // [B][A]
  int* n_cand;             // [B]
  int* n_seg;              // [B]
  unsigned char* flags;    // [B][A]  bit0 = suppressed, bit1 = kept
  unsigned long long* keys_g;  // [B][A_pow2] or null (global-memory sort for very large A)
};

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

static size_t carve(NmsWork* w, void* base, int batch, int anchors) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return reinterpret_cast<unsigned char*>(base) + o;
  };
  const size_t ba = static_cast<size_t>(batch) * anchors;
  unsigned char* p;
  p = take(ba * 4); if (w) w->order = reinterpret_cast<int*>(p);
  p = take(ba * 4); if (w) w->seg_start = reinterpret_cast<int*>(p);
  p = take(ba * 4); if (w) w->seg_len = reinterpret_cast<int*>(p);
  p = take(static_cast<size_t>(batch) * 4); if (w) w->n_cand = reinterpret_cast<int*>(p);
  p = take(static_cast<size_t>(batch) * 4); if (w) w->n_seg = reinterpret_cast<int*>(p);
  p = take(ba); if (w) w->flags = p;
  const int ap2 = next_pow2(anchors);
  if (ap2 > kSortSmemKeys) {
    p = take(static_cast<size_t>(batch) * ap2 * 8);
    if (w) w->keys_g = reinterpret_cast<unsigned long long*>(p);
  } else if (w) {
    w->keys_g = nullptr;
  }
  return off;
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nms_prepare_dense_kernel(float* __restrict__ pred, long long rows, int nc,
                                                                float* __restrict__ cand_box,
                                                                float* __restrict__ cand_conf,
                                                                int* __restrict__ cand_cls) {
  const long long row = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* r = pred + row * (4 + nc);
  // class max, first maximum wins (torch.max on CPU)
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int c = lane; c < nc; c += 32) {
    const float v = r[4 + c];
    if (v > best || bi == 0x7fffffff) { best = v; bi = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > best || (ov == best && oi < bi))) { best = ov; bi = oi; }
  }
  const float bx = r[0], by = r[1], bw = r[2], bh = r[3];
  __syncwarp();
  if (lane == 0) {
    const float x1 = bx - bw / 2.0f, y1 = by - bh / 2.0f, x2 = bx + bw / 2.0f, y2 = by + bh / 2.0f;
    r[0] = x1; r[1] = y1; r[2] = x2; r[3] = y2;
    reinterpret_cast<float4*>(cand_box)[row] = make_float4(x1, y1, x2, y2);
    cand_conf[row] = best;
    cand_cls[row] = bi;
  }
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int score_desc_bits(float s) {
  unsigned int u = __float_as_uint(s);
  u ^= (u >> 31) ? 0xffffffffu : 0x80000000u;  // ascending-sortable
  return ~u;                                   // descending
}

__global__ void __launch_bounds__(kSortThreads) nms_sort_kernel(const float* __restrict__ cand_conf,
                                                                const int* __restrict__ cand_cls, int anchors,
                                                                float conf_thres, NmsWork wk, int a_pow2, const TimelineTag tl) {
  const unsigned long long tl_t0 = (tl.buf != nullptr && threadIdx.x == 0) ? global_timer_ns() : 0ull;
  extern __shared__ unsigned long long sort_smem[];
  __shared__ int s_count;
  const int b = blockIdx.x;
  unsigned long long* keys = wk.keys_g ? wk.keys_g + static_cast<size_t>(b) * a_pow2 : sort_smem;
  const float* conf = cand_conf + static_cast<size_t>(b) * anchors;
  const int* cls = cand_cls + static_cast<size_t>(b) * anchors;
  unsigned char* flags = wk.flags + static_cast<size_t>(b) * anchors;

  if (threadIdx.x == 0) s_count = 0;
  for (int i = threadIdx.x; i < a_pow2; i += kSortThreads) keys[i] = ~0ull;
  for (int i = threadIdx.x; i < anchors; i += kSortThreads) flags[i] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < anchors; i += kSortThreads) {
    const float s = conf[i];
    // class ids outside [0, 4096) do not fit the key's 12 class bits (they would reorder the segments silently):
    // such a row -- e.g. the 0x7fffffff of an all-NaN score row -- is never a candidate
    if (s >= conf_thres && static_cast<unsigned>(cls[i]) < 4096u) {
      const int pos = atomicAdd(&s_count, 1);
      keys[pos] = (static_cast<unsigned long long>(cls[i]) << (kIdxBits + kScoreBits)) |
                  (static_cast<unsigned long long>(score_desc_bits(s)) << kIdxBits) | static_cast<unsigned long long>(i);
    }
  }
  __syncthreads();
  const int n = s_count;
  __syncthreads();  // everyone has read the count before it is reused below
  int n_pad = 1;
  while (n_pad < n) n_pad <<= 1;
  // bitonic sort, ascending
  for (int k = 2; k <= n_pad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (n_pad >> 1); t += kSortThreads) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int l = i | j;
        const unsigned long long a = keys[i], c = keys[l];
        const bool up = (i & k) == 0;
        if ((a > c) == up) { keys[i] = c; keys[l] = a; }
      }
      __syncthreads();
    }
  }
  int* order = wk.order + static_cast<size_t>(b) * anchors;
  int* seg_start = wk.seg_start + static_cast<size_t>(b) * anchors;
  int* seg_len = wk.seg_len + static_cast<size_t>(b) * anchors;
  if (threadIdx.x == 0) { s_count = 0; wk.n_cand[b] = n; }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += kSortThreads) {
    const unsigned long long key = keys[i];
    order[i] = static_cast<int>(key & kIdxMask);
    const unsigned long long c = key >> (kIdxBits + kScoreBits);
    if (i == 0 || (keys[i - 1] >> (kIdxBits + kScoreBits)) != c) {
      // end of this class segment = first key of a larger class (binary search in the sorted keys)
      const unsigned long long bound = (c + 1) << (kIdxBits + kScoreBits);
      int lo = i + 1, hi = n;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (keys[mid] < bound) lo = mid + 1; else hi = mid;
      }
      const int slot = atomicAdd(&s_count, 1);
      seg_start[slot] = i;
      seg_len[slot] = lo - i;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) wk.n_seg[b] = s_count;
  if (threadIdx.x == 0) timeline_write(tl, tl_t0);
}

// ------------------------------------------------------------------------------------------------
// torchvision nms_kernel_impl arithmetic: inter / (area_i + area_j - inter) > thr
__device__ __forceinline__ bool iou_gt(const float4 a, const float area_a, const float4 b, const float area_b,
                                       const float thr) {
  const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
  const float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
  const float w = fmaxf(0.0f, xx2 - xx1), h = fmaxf(0.0f, yy2 - yy1);
  const float inter = w * h;
  // disjoint boxes: inter == 0 -> iou is 0 (or NaN for two empty boxes), never > a non-negative threshold.  Skipping
  // the IEEE division for them leaves every comparison result unchanged.
  if (inter == 0.0f && thr >= 0.0f) return false;
  const float ovr = inter / (area_a + area_b - inter);
  return ovr > thr;
}

// Greedy NMS of one class segment processes boxes in score order; a box's fate is known once it has been compared with
// every EARLIER KEPT box.
//
// Segments of up to kWarpSegMax boxes (the common case: ~80 classes share a few thousand candidates) are handled by ONE
// WARP each, 32 boxes at a time: a tile is first swept against the kept boxes of earlier tiles (a per-warp list in
// shared memory), then resolved within itself (per-lane 32-bit IoU masks + a shuffle scan).  No block barriers and no
// flag re-reads: many independent warps hide one another's load latency.
// Larger segments fall back to the CTA-cooperative rounds below: warp 0 gathers the next <= 32 boxes that are still
// alive (dead ones are skipped for good), resolves them among themselves and publishes the kept ones; then the whole
// CTA sweeps the not-yet-visited tail of the segment against those kept boxes.
// Work is O(kept x segment) instead of O(segment^2); the result is exactly the sequential greedy one.
constexpr int kWarpSegMax = 128;     // boxes per warp-handled segment
constexpr int kCtaSegSmem = 2048;     // boxes of a CTA-handled segment staged in shared memory (32 KB + flags)

__device__ __forceinline__ void nms_warp_segment(const float4* __restrict__ boxes, const int* __restrict__ ord,
                                                 unsigned char* __restrict__ flags, int len, float iou_thr, float4* kept_list,
                                                 int lane) {
  const unsigned full = 0xffffffffu;
  int nk = 0;
  for (int t0 = 0; t0 < len; t0 += 32) {
    const int p = t0 + lane;
    const bool has = p < len;
    float4 bi = make_float4(0.f, 0.f, 0.f, 0.f);
    if (has) bi = boxes[ord[p]];
    const float area_i = (bi.z - bi.x) * (bi.w - bi.y);
    bool alive = has;
    for (int j = 0; j < nk && alive; ++j) {          // kept boxes of earlier tiles
      const float4 bj = kept_list[j];
      if (iou_gt(bj, (bj.z - bj.x) * (bj.w - bj.y), bi, area_i, iou_thr)) alive = false;
    }
    unsigned alive_bits = __ballot_sync(full, alive);
    const int cnt = min(32, len - t0);
    unsigned mask = 0;                                // later in-tile boxes this box suppresses
    for (int j = 1; j < cnt; ++j) {
      float4 bj;
      bj.x = __shfl_sync(full, bi.x, j);
      bj.y = __shfl_sync(full, bi.y, j);
      bj.z = __shfl_sync(full, bi.z, j);
      bj.w = __shfl_sync(full, bi.w, j);
      const float area_j = __shfl_sync(full, area_i, j);
      if (j > lane && iou_gt(bi, area_i, bj, area_j, iou_thr)) mask |= 1u << j;
    }
    for (int j = 0; j < cnt; ++j) {
      const unsigned mj = __shfl_sync(full, mask, j);
      if ((alive_bits >> j) & 1u) alive_bits &= ~mj;
    }
    const bool kept = has && ((alive_bits >> lane) & 1u);
    if (has) flags[p] = kept ? 2 : 1;
    if (t0 + 32 < len) {                              // only later tiles read the list
      if (kept) kept_list[nk + __popc(alive_bits & ((1u << lane) - 1u))] = bi;
      nk += __popc(alive_bits);
      __syncwarp();
    }
  }
}

// One segment, whole CTA.  SMEM: boxes / flags are the staged shared-memory copies (index = position in the segment);
// otherwise boxes are gathered through ord[] and the flags live in global memory.  flags: bit0 = suppressed, bit1 = kept.
template <bool SMEM>
__device__ __forceinline__ void nms_cta_segment(const float4* boxes, const int* ord, unsigned char* flags, int len, float iou_thr,
                                                float4* s_kept, int* s_tilepos, int* s_ctl) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  auto box_at = [&](int k) { return SMEM ? boxes[k] : boxes[ord[k]]; };
  if (threadIdx.x == 0) s_ctl[0] = 0;
  __syncthreads();
  while (true) {
    if (warp == 0) {
      // ---- gather the next <= 32 alive boxes in score order
      int pos = s_ctl[0], cnt = 0;
      while (cnt < 32 && pos < len) {
        const int pp = pos + lane;
        const bool al = pp < len && (flags[pp] & 1) == 0;
        const unsigned bal = __ballot_sync(full, al);
        const int avail = __popc(bal);
        const int take = min(avail, 32 - cnt);
        const int rank = __popc(bal & ((1u << lane) - 1u));
        if (al && rank < take) s_tilepos[cnt + rank] = pp;
        if (take < avail) {
          const unsigned last = __ballot_sync(full, al && rank == take - 1);
          pos += __ffs(last);          // one past the last box taken
        } else {
          pos += 32;
        }
        cnt += take;
      }
      __syncwarp();
      // ---- resolve the tile: lane i owns the i-th gathered box (all are alive on entry)
      const bool has = lane < cnt;
      float4 bi = make_float4(0.f, 0.f, 0.f, 0.f);
      int my = 0;
      if (has) {
        my = s_tilepos[lane];
        bi = box_at(my);
      }
      const float area_i = (bi.z - bi.x) * (bi.w - bi.y);
      unsigned mask = 0;   // later in-tile boxes this box suppresses
      for (int j = 1; j < cnt; ++j) {
        float4 bj;
        bj.x = __shfl_sync(full, bi.x, j);
        bj.y = __shfl_sync(full, bi.y, j);
        bj.z = __shfl_sync(full, bi.z, j);
        bj.w = __shfl_sync(full, bi.w, j);
        const float area_j = __shfl_sync(full, area_i, j);
        if (j > lane && iou_gt(bi, area_i, bj, area_j, iou_thr)) mask |= 1u << j;
      }
      unsigned alive_bits = cnt >= 32 ? full : ((1u << cnt) - 1u);
      for (int j = 0; j < cnt; ++j) {
        const unsigned mj = __shfl_sync(full, mask, j);
        if ((alive_bits >> j) & 1u) alive_bits &= ~mj;
      }
      const bool kept = has && ((alive_bits >> lane) & 1u);
      if (has) flags[my] |= kept ? 2 : 1;
      if (kept) s_kept[__popc(alive_bits & ((1u << lane) - 1u))] = bi;
      if (lane == 0) {
        s_ctl[0] = pos < len ? pos : len;
        s_ctl[1] = cnt;
        s_ctl[2] = __popc(alive_bits);
      }
    }
    __syncthreads();
    const int cnt = s_ctl[1], nk = s_ctl[2], pos = s_ctl[0];
    if (cnt == 0) break;
    // ---- kept boxes of this round suppress the unvisited tail
    for (int k = pos + threadIdx.x; k < len; k += kSegWarps * 32) {
      if (flags[k] & 1) continue;
      const float4 bk = box_at(k);
      const float area_k = (bk.z - bk.x) * (bk.w - bk.y);
      for (int j = 0; j < nk; ++j) {
        const float4 bj = s_kept[j];
        if (iou_gt(bj, (bj.z - bj.x) * (bj.w - bj.y), bk, area_k, iou_thr)) {
          flags[k] |= 1;
          break;
        }
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kSegWarps * 32) nms_segment_kernel(const float4* __restrict__ cand_box, int anchors,
                                                                     float iou_thr, NmsWork wk, const TimelineTag tl) {
  const unsigned long long tl_t0 = (tl.buf != nullptr && threadIdx.x == 0) ? global_timer_ns() : 0ull;
  // pass 1 uses the union as per-warp kept lists [kSegWarps][kWarpSegMax], pass 2 as staged boxes + flags of one segment
  __shared__ __align__(16) unsigned char s_union[kCtaSegSmem * (sizeof(float4) + 1)];
  __shared__ float4 s_kept[32];
  __shared__ int s_tilepos[32];
  __shared__ int s_ctl[4];
  static_assert(kSegWarps * kWarpSegMax <= kCtaSegSmem, "per-warp kept lists must fit the union");
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nseg = wk.n_seg[b];
  const float4* boxes = cand_box + static_cast<size_t>(b) * anchors;
  const int* order = wk.order + static_cast<size_t>(b) * anchors;
  unsigned char* flags_img = wk.flags + static_cast<size_t>(b) * anchors;

  // ---- pass 1: one warp per small segment
  for (int seg = blockIdx.x * kSegWarps + warp; seg < nseg; seg += gridDim.x * kSegWarps) {
    const int s0 = wk.seg_start[b * static_cast<size_t>(anchors) + seg];
    const int len = wk.seg_len[b * static_cast<size_t>(anchors) + seg];
    if (len <= kWarpSegMax)
      nms_warp_segment(boxes, order + s0, flags_img + s0, len, iou_thr, reinterpret_cast<float4*>(s_union) + warp * kWarpSegMax, lane);
  }
  // ---- pass 2: CTA-cooperative rounds for the large segments (boxes staged in shared memory when they fit)
  __syncthreads();   // pass 1 is done with the shared-memory union
  for (int seg = blockIdx.x; seg < nseg; seg += gridDim.x) {
    const int s0 = wk.seg_start[b * static_cast<size_t>(anchors) + seg];
    const int len = wk.seg_len[b * static_cast<size_t>(anchors) + seg];
    if (len <= kWarpSegMax) continue;        // block-uniform
    if (len <= kCtaSegSmem) {
      float4* sb = reinterpret_cast<float4*>(s_union);
      unsigned char* sf = s_union + kCtaSegSmem * sizeof(float4);
      for (int i = threadIdx.x; i < len; i += kSegWarps * 32) {
        sb[i] = boxes[order[s0 + i]];
        sf[i] = 0;
      }
      __syncthreads();
      nms_cta_segment<true>(sb, nullptr, sf, len, iou_thr, s_kept, s_tilepos, s_ctl);
      for (int i = threadIdx.x; i < len; i += kSegWarps * 32) flags_img[s0 + i] = sf[i];
      __syncthreads();
    } else {
      nms_cta_segment<false>(boxes, order + s0, flags_img + s0, len, iou_thr, s_kept, s_tilepos, s_ctl);
    }
  }
  if (threadIdx.x == 0) timeline_write(tl, tl_t0);
}

// ------------------------------------------------------------------------------------------------
// correct_boxes (utils/bbox_utils.py:84-117) on the kept rows, in numpy's dtype flow (this file is compiled with
// -fmad=false): centres / sizes in float32 ((x1+x2)/2, x2-x1, :179); with letterbox the centre goes to float64
// ((yx - offset) * scale) while the size is scaled in float64 and stored back to float32 (`box_hw *= scale` is in place on
// a float32 view); mins / maxes in float64, times the image shape, stored as float32.  Without letterbox everything
// stays float32 and only the final in-place `boxes *= image_shape` passes through float64.
// params [B][6] f64: offset_y, offset_x, scale_y, scale_x, image_h, image_w (computed on the host with the reference's
// own expressions).  rows [x1, y1, x2, y2, conf, cls] -> [y1, x1, y2, x2, conf, cls] in image pixels.
__global__ void __launch_bounds__(128) correct_boxes_kernel(const float* __restrict__ dets, const int* __restrict__ keep_count,
                                                            int anchors, const double* __restrict__ params, int letterbox,
                                                            float* __restrict__ out) {
  const int b = blockIdx.y;
  const int n = keep_count[b];
  const double* q = params + b * 6;
  const double off[2] = {q[0], q[1]}, sc[2] = {q[2], q[3]}, dim[2] = {q[4], q[5]};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float* d = dets + (static_cast<size_t>(b) * anchors + i) * 6;
    float* o = out + (static_cast<size_t>(b) * anchors + i) * 6;
    const float x1 = d[0], y1 = d[1], x2 = d[2], y2 = d[3];
    const float c32[2] = {(y1 + y2) / 2.0f, (x1 + x2) / 2.0f};    // box_yx
    const float s32[2] = {y2 - y1, x2 - x1};                      // box_hw
    float r[4];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      if (letterbox) {
        const double c = (static_cast<double>(c32[a]) - off[a]) * sc[a];
        const float hw = static_cast<float>(static_cast<double>(s32[a]) * sc[a]);
        const double half = static_cast<double>(hw / 2.0f);
        r[a] = static_cast<float>((c - half) * dim[a]);
        r[a + 2] = static_cast<float>((c + half) * dim[a]);
      } else {
        const float half = s32[a] / 2.0f;
        r[a] = static_cast<float>(static_cast<double>(c32[a] - half) * dim[a]);
        r[a + 2] = static_cast<float>(static_cast<double>(c32[a] + half) * dim[a]);
      }
    }
    const float conf = d[4], cls = d[5];
    o[0] = r[0]; o[1] = r[1]; o[2] = r[2]; o[3] = r[3];
    o[4] = conf; o[5] = cls;
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nms_compact_kernel(const float4* __restrict__ cand_box,
                                                          const float* __restrict__ cand_conf,
                                                          const int* __restrict__ cand_cls, int anchors, NmsWork wk,
                                                          int* __restrict__ keep_idx, int* __restrict__ keep_count,
                                                          float* __restrict__ dets) {
  __shared__ int s_warp[8];
  __shared__ int s_base;
  const int b = blockIdx.x;
  const int n = wk.n_cand[b];
  const int* order = wk.order + static_cast<size_t>(b) * anchors;
  const unsigned char* flags = wk.flags + static_cast<size_t>(b) * anchors;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += 256) {
    const int i = i0 + threadIdx.x;
    const bool kept = i < n && (flags[i] & 2);
    const unsigned bal = __ballot_sync(0xffffffffu, kept);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    if (kept) {
      const int pos = off + __popc(bal & ((1u << lane) - 1u));
      const int a = order[i];
      const size_t g = static_cast<size_t>(b) * anchors;
      keep_idx[g + pos] = a;
      if (dets) {
        const float4 bx = cand_box[g + a];
        float* d = dets + (g + pos) * 6;
        d[0] = bx.x; d[1] = bx.y; d[2] = bx.z; d[3] = bx.w;
        d[4] = cand_conf[g + a];
        d[5] = static_cast<float>(cand_cls[g + a]);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < 8; ++w) tot += s_warp[w];
      s_base += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) keep_count[b] = s_base;
}


// ------------------------------------------------------------------------------------------------
// Result packing: the kept rows of a batch, compacted image after image, so that ONE small
// device-to-host copy carries every detection -- and, with max_boxes > 0, the top-`max_boxes` selection of the reference's
// detect loop, utils/callbacks.py:159-166 (`top_100 = np.argsort(top_conf)[::-1][:self.max_boxes]`).  Tie order is defined
// here as score descending, then kept order (the reference's numpy argsort is an unstable sort and defines none).
// One CTA per image: offset = sum of the earlier images' output counts; keys (~score | row) bitonic-sorted in shared
// memory (global workspace beyond kSortSmemKeys rows).
__global__ void __launch_bounds__(kSortThreads) pack_detections_kernel(const float* __restrict__ rows, const int* __restrict__ keep_count,
                                                                       int anchors, int max_boxes, int* __restrict__ offsets,
                                                                       float* __restrict__ out, unsigned long long* __restrict__ keys_g,
                                                                       int a_pow2) {
  extern __shared__ unsigned long long sort_smem[];
  __shared__ int s_part[kSortThreads / 32];
  __shared__ int s_off;
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = keep_count[b];
  const int k = max_boxes > 0 ? min(n, max_boxes) : n;
  int part = 0;
  for (int i = threadIdx.x; i < b; i += kSortThreads) part += max_boxes > 0 ? min(keep_count[i], max_boxes) : keep_count[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (lane == 0) s_part[warp] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    int off = 0;
    for (int w = 0; w < kSortThreads / 32; ++w) off += s_part[w];
    s_off = off;
    offsets[b] = off;
    if (b == static_cast<int>(gridDim.x) - 1) offsets[b + 1] = off + k;
  }
  __syncthreads();
  const float* src = rows + static_cast<size_t>(b) * anchors * 6;
  float* dst = out + static_cast<size_t>(s_off) * 6;
  if (max_boxes <= 0) {                       // plain compaction in the NMS output order
    for (int i = threadIdx.x; i < n * 6; i += kSortThreads) dst[i] = src[i];
    return;
  }
  unsigned long long* keys = keys_g ? keys_g + static_cast<size_t>(b) * a_pow2 : sort_smem;
  int n_pad = 1;
  while (n_pad < n) n_pad <<= 1;
  for (int i = threadIdx.x; i < n_pad; i += kSortThreads)
    keys[i] = i < n ? (static_cast<unsigned long long>(score_desc_bits(src[i * 6 + 4])) << 32) | static_cast<unsigned>(i) : ~0ull;
  __syncthreads();
  for (int kk = 2; kk <= n_pad; kk <<= 1) {
    for (int j = kk >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (n_pad >> 1); t += kSortThreads) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int l = i | j;
        const unsigned long long a = keys[i], c = keys[l];
        const bool up = (i & kk) == 0;
        if ((a > c) == up) { keys[i] = c; keys[l] = a; }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < k * 6; i += kSortThreads) {
    const int r = i / 6, c = i - r * 6;
    dst[i] = src[static_cast<size_t>(keys[r] & 0xffffffffull) * 6 + c];
  }
}

}  // namespace tod

using namespace tod;

extern "C" int tod_nms_prepare_dense(float* d_prediction, int32_t batch, int32_t anchors, int32_t nc,
                                     float* d_cand_box, float* d_cand_conf, int32_t* d_cand_cls, void* stream) {
  TOD_CHECK_ARG(d_prediction && d_cand_box && d_cand_conf && d_cand_cls, "nms_prepare_dense: null pointer");
  TOD_CHECK_ARG(batch > 0 && anchors > 0 && nc > 0, "nms_prepare_dense: bad shape");
  TOD_CHECK_ARG(nc <= 4096, "nms_prepare_dense: nc %d exceeds the 4096 classes the sort key holds", nc);
  TOD_CHECK_ARG((reinterpret_cast<uintptr_t>(d_cand_box) & 15) == 0, "nms_prepare_dense: cand_box not 16-byte aligned");
  const long long rows = static_cast<long long>(batch) * anchors;
  const long long blocks = (rows * 32 + 255) / 256;
  TOD_CHECK_ARG(blocks < (1ll << 31), "nms_prepare_dense: too many rows");
  nms_prepare_dense_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_prediction, rows, nc, d_cand_box, d_cand_conf, d_cand_cls);
  TOD_CHECK_LAUNCH("nms_prepare_dense_kernel launch");
  return TOD_OK;
}

extern "C" int64_t tod_nms_workspace_bytes(int32_t batch, int32_t anchors) {
  if (batch <= 0 || anchors <= 0) return 0;
  return static_cast<int64_t>(carve(nullptr, nullptr, batch, anchors));
}

extern "C" int tod_nms(const float* d_cand_box, const float* d_cand_conf, const int32_t* d_cand_cls, int32_t batch,
                       int32_t anchors, float conf_thres, double iou_thres, void* d_work, int64_t work_bytes,
                       int32_t* d_keep_idx, int32_t* d_keep_count, float* d_dets, void* stream) {
  TOD_CHECK_ARG(d_cand_box && d_cand_conf && d_cand_cls && d_work && d_keep_idx && d_keep_count, "nms: null pointer");
  TOD_CHECK_ARG(batch > 0 && batch <= 65535 && anchors > 0 && anchors < (1 << kIdxBits), "nms: bad batch %d / anchors %d",
                batch, anchors);
  TOD_CHECK_ARG((reinterpret_cast<uintptr_t>(d_cand_box) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_work) & 255) == 0,
                "nms: cand_box must be 16-byte and workspace 256-byte aligned");
  NmsWork wk;
  const size_t need = carve(&wk, d_work, batch, anchors);
  TOD_CHECK_ARG(work_bytes >= static_cast<int64_t>(need), "nms: workspace %lld < %lld bytes", (long long)work_bytes,
                (long long)need);
  // torchvision's CPU kernel compares (double)iou > thr; for a float iou that is iou > largest float <= thr
  float thr_f = static_cast<float>(iou_thres);
  if (static_cast<double>(thr_f) > iou_thres) thr_f = nextafterf(thr_f, -INFINITY);
  auto st = static_cast<cudaStream_t>(stream);
  const int ap2 = next_pow2(anchors);
  const size_t sort_smem = wk.keys_g ? 0 : static_cast<size_t>(ap2) * 8;
  static PerDeviceOnce attr_once;   // the attribute is per device
  if (attr_once.needed()) {
    int rc = check_cuda(cudaFuncSetAttribute(nms_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kSortSmemKeys * 8),
                        "cudaFuncSetAttribute(nms_sort)");
    if (rc != TOD_OK) return rc;
    attr_once.done();
  }
  nms_sort_kernel<<<batch, kSortThreads, sort_smem, st>>>(d_cand_conf, d_cand_cls, anchors, conf_thres, wk, ap2, timeline_tag("nms sort"));
  TOD_CHECK_LAUNCH("nms_sort_kernel launch");
  nms_segment_kernel<<<dim3(kSegCtasPerImage, batch), kSegWarps * 32, 0, st>>>(
      reinterpret_cast<const float4*>(d_cand_box), anchors, thr_f, wk, timeline_tag("nms segments"));
  TOD_CHECK_LAUNCH("nms_segment_kernel launch");
  nms_compact_kernel<<<batch, 256, 0, st>>>(reinterpret_cast<const float4*>(d_cand_box), d_cand_conf, d_cand_cls, anchors,
                                            wk, d_keep_idx, d_keep_count, d_dets);
  TOD_CHECK_LAUNCH("nms_compact_kernel launch");
  return TOD_OK;
}

extern "C" int tod_correct_boxes(const float* d_dets, const int32_t* d_keep_count, int32_t batch, int32_t anchors,
                                 const double* d_params, int32_t letterbox, float* d_rows, void* stream) {
  TOD_CHECK_ARG(d_dets != nullptr && d_keep_count != nullptr && d_params != nullptr && d_rows != nullptr, "correct_boxes: null pointer");
  TOD_CHECK_ARG(batch > 0 && batch <= 65535 && anchors > 0, "correct_boxes: bad sizes");
  correct_boxes_kernel<<<dim3(4, batch), 128, 0, static_cast<cudaStream_t>(stream)>>>(d_dets, d_keep_count, anchors, d_params,
                                                                                      letterbox, d_rows);
  TOD_CHECK_LAUNCH("correct_boxes_kernel launch");
  return TOD_OK;
}

extern "C" int64_t tod_pack_workspace_bytes(int32_t batch, int32_t anchors) {
  if (batch <= 0 || anchors <= 0) return 0;
  const int ap2 = next_pow2(anchors);
  return ap2 > kSortSmemKeys ? static_cast<int64_t>(batch) * ap2 * 8 : 0;
}

extern "C" int tod_pack_detections(const float* d_rows, const int32_t* d_keep_count, int32_t batch, int32_t anchors,
                                   int32_t max_boxes, int32_t* d_offsets, float* d_packed, void* d_work, int64_t work_bytes,
                                   void* stream) {
  TOD_CHECK_ARG(d_rows && d_keep_count && d_offsets && d_packed, "pack_detections: null pointer");
  TOD_CHECK_ARG(batch > 0 && batch <= 65535 && anchors > 0, "pack_detections: bad sizes");
  const int ap2 = next_pow2(anchors);
  const int64_t need = max_boxes > 0 ? tod_pack_workspace_bytes(batch, anchors) : 0;
  TOD_CHECK_ARG(need == 0 || (d_work != nullptr && work_bytes >= need && (reinterpret_cast<uintptr_t>(d_work) & 7) == 0),
                "pack_detections: workspace %lld < %lld bytes", (long long)work_bytes, (long long)need);
  static PerDeviceOnce attr_once;
  if (attr_once.needed()) {
    int rc = check_cuda(cudaFuncSetAttribute(pack_detections_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSortSmemKeys * 8),
                        "cudaFuncSetAttribute(pack_detections)");
    if (rc != TOD_OK) return rc;
    attr_once.done();
  }
  const size_t smem = (max_boxes > 0 && need == 0) ? static_cast<size_t>(ap2) * 8 : 0;
  pack_detections_kernel<<<batch, kSortThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      d_rows, d_keep_count, anchors, max_boxes, d_offsets, d_packed,
      need ? reinterpret_cast<unsigned long long*>(d_work) : nullptr, ap2);
  TOD_CHECK_LAUNCH("pack_detections_kernel launch");
  return TOD_OK;
}

// Fused head decode: DFL softmax-expectation + anchor/stride box decode + class sigmoid + class max,
// emitting (any subset of) the Head eval tensor, the decode_box tensor and NMS-ready candidates.
// Replaces Head.forward eval branch (model/head.py:53-61), DFL.forward (model/blocks.py:154-157),
// make_anchors (utils/bbox_utils.py:14-37), dist2bbox (:39-58), DecodeBox.decode_box (:66-82) and the
// corner conversion + class max at the top of non_max_suppression (:144-153).
// Arithmetic follows SURVEY.md Appendix B step by step in fp32 (true divisions, no FMA contraction:
// this file is compiled with -fmad=false).  HBM-bound: reads (64+nc) f32 per anchor, writes up to
// 2*(4+nc)+6 f32 per anchor; every global access is a coalesced row segment.
#include "decode_math.cuh"

namespace tod {

constexpr int kDecThreads = 256;
constexpr int kDecAnchors = 64;  // anchors per CTA (4 threads per anchor)

struct DecodeParams {
  const float* raw[3];
  int h[3], w[3];
  float stride[3];
  int tiles[3];       // CTAs per level
  int level_off[3];   // first anchor index of the level
  int raw_pitch, batch, nc, anchors;
  float in_w, in_h;
  float* head_out;
  float* decoded;
  float* cand_box;
  float* cand_conf;
  int* cand_cls;
};

// Four threads per anchor row (64 + nc f32 logits), 64 anchors per CTA.  Thread (a, j): DFL of box side j (16 bins,
// 64 contiguous bytes) and a contiguous quarter of the class logits, so the four threads of an anchor read one
// contiguous run.  FULL: every class score is materialised (Head tensor / decode_box tensor requested).  Otherwise only
// the NMS candidate is needed: the class maximum is found on the LOGITS (each thread keeps its two largest) and the
// sigmoid is evaluated only for logits within a window of the maximum wide enough to contain every float32 tie of the
// scores (the reference takes max over the sigmoid outputs and keeps the lowest class among equal scores); where the
// sigmoid saturates the window is everything.  Both modes produce bit-identical candidates.
template <bool FULL>
__global__ void __launch_bounds__(kDecThreads) head_decode_kernel(const DecodeParams p) {
  extern __shared__ float dec_smem[];
  const int spitch = p.nc + 1;
  float* s_score = dec_smem;                           // [64][nc + 1]     (FULL only)
  float* s_box = dec_smem + kDecAnchors * spitch;      // [64][4]  xywh in input pixels

  int lvl = 0, tile = blockIdx.x;
  if (tile >= p.tiles[0]) { tile -= p.tiles[0]; lvl = 1; }
  if (lvl == 1 && tile >= p.tiles[1]) { tile -= p.tiles[1]; lvl = 2; }
  const int b = blockIdx.y;
  const int lw = p.w[lvl];
  const int la = p.h[lvl] * lw;             // anchors in this level
  const int a0 = tile * kDecAnchors;        // first anchor (level-local) of this CTA
  const int na = min(kDecAnchors, la - a0);
  const float* raw = p.raw[lvl] + (static_cast<size_t>(b) * la + a0) * p.raw_pitch;
  const float stride = p.stride[lvl];
  const unsigned full = 0xffffffffu;
  const int ag0 = p.level_off[lvl] + a0;    // global anchor index of the CTA's first anchor
  const float ninf = -INFINITY;

  const int a = threadIdx.x >> 2, side = threadIdx.x & 3;
  const bool valid = a < na;
  const float4* row = reinterpret_cast<const float4*>(raw + static_cast<size_t>(valid ? a : 0) * p.raw_pitch);
  // class chunks (4 classes each) of this thread: [k0, k1) of the ceil(nc / 4) chunks after the 16 box chunks
  const int kc = (p.nc + 3) >> 2, per = (kc + 3) >> 2;
  const int k0 = min(side * per, kc), k1 = min(k0 + per, kc);

  // ---- DFL (blocks.py:154-157): softmax over the 16 bins of this side, expectation with arange(16)
  float dist = 0.f;
  if (valid) {
    float l[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 v = __ldg(row + side * 4 + i);
      l[4 * i] = v.x; l[4 * i + 1] = v.y; l[4 * i + 2] = v.z; l[4 * i + 3] = v.w;
    }
    dist = dfl_side(l);
  }
  // ---- classes
  float best = -1.0f;
  int bi = 0x7fffffff;
  if (FULL) {
    if (valid)
      for (int k = k0; k < k1; ++k) {
        const float4 v = __ldg(row + 16 + k);
        const float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = 4 * k + e;
          if (c < p.nc) {
            const float sc = sigmoid_ref(xs[e]);
            s_score[a * spitch + c] = sc;
            better(best, bi, sc, c);
          }
        }
      }
  } else {
    // the two largest logits of this thread (first occurrence wins ties, like the ascending-class scan of the scores)
    LogitTop2 top;
    if (valid)
      for (int k = k0; k < k1; ++k) {
        const float4 v = __ldg(row + 16 + k);
        const float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (4 * k + e < p.nc) top.add(xs[e], 4 * k + e);
      }
    const float x1 = top.x1, x2 = top.x2;
    const int c1 = top.c1;
    float m = x1;
    m = fmaxf(m, __shfl_xor_sync(full, m, 1));
    m = fmaxf(m, __shfl_xor_sync(full, m, 2));
    const float thr_logit = tie_window_threshold(m);
    if (valid) {
      if (x2 >= thr_logit) {                 // rare: more than one of this thread's logits in the window -> rescan
        for (int k = k0; k < k1; ++k) {
          const float4 v = __ldg(row + 16 + k);
          const float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = 4 * k + e;
            if (c < p.nc && xs[e] >= thr_logit) better(best, bi, sigmoid_ref(xs[e]), c);
          }
        }
      } else if (x1 >= thr_logit) {
        best = sigmoid_ref(x1);
        bi = c1;
      }
    }
  }
  // ---- combine the four threads of the anchor (lanes 4q .. 4q+3)
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1) {
    const float s2 = __shfl_xor_sync(full, best, o);
    const int c2 = __shfl_xor_sync(full, bi, o);
    better(best, bi, s2, c2);
  }
  const int base = (threadIdx.x & 31) & ~3;
  const float dl = __shfl_sync(full, dist, base + 0);
  const float dt = __shfl_sync(full, dist, base + 1);
  const float dr = __shfl_sync(full, dist, base + 2);
  const float db = __shfl_sync(full, dist, base + 3);
  if (side == 0 && valid) {
    const int ai = a0 + a;
    const int gy = ai / lw, gx = ai - gy * lw;
    const float4 bpx = box_xywh_px(dl, dt, dr, db, gx, gy, stride);
    const float bx = bpx.x, by = bpx.y, bw = bpx.z, bh = bpx.w;
    if (FULL) {
      s_box[a * 4 + 0] = bx;
      s_box[a * 4 + 1] = by;
      s_box[a * 4 + 2] = bw;
      s_box[a * 4 + 3] = bh;
    }
    // NMS candidate: corners of the normalised xywh (bbox_utils.py:144-148), class max (:153)
    if (p.cand_conf) {
      const size_t g = static_cast<size_t>(b) * p.anchors + ag0 + a;
      p.cand_conf[g] = best;
      p.cand_cls[g] = bi;
      reinterpret_cast<float4*>(p.cand_box)[g] = box_corners_norm(bpx, p.in_w, p.in_h);
    }
  }
  if (!FULL) return;
  __syncthreads();

  const int no = 4 + p.nc;
  // ---- Head eval tensor (B, 4+nc, A): contiguous along anchors
  if (p.head_out) {
    float* o = p.head_out + static_cast<size_t>(b) * no * p.anchors + ag0;
    for (int i = threadIdx.x; i < no * kDecAnchors; i += kDecThreads) {
      const int ch = i / kDecAnchors, a = i - ch * kDecAnchors;
      if (a < na) o[static_cast<size_t>(ch) * p.anchors + a] = ch < 4 ? s_box[a * 4 + ch] : s_score[a * spitch + ch - 4];
    }
  }
  // ---- decode_box tensor (B, A, 4+nc): contiguous rows, xywh / (W,H,W,H)
  if (p.decoded) {
    float* o = p.decoded + (static_cast<size_t>(b) * p.anchors + ag0) * no;
    for (int i = threadIdx.x; i < na * no; i += kDecThreads) {
      const int a = i / no, ch = i - a * no;
      float v;
      if (ch < 4) v = s_box[a * 4 + ch] / ((ch & 1) ? p.in_h : p.in_w);
      else v = s_score[a * spitch + ch - 4];
      o[i] = v;
    }
  }
}

// DecodeBox.decode_box applied to an existing Head eval tensor (utils/bbox_utils.py:77-82 on the head tensor,
// SURVEY F7): (B, 4+nc, A) -> (B, A, 4+nc) with xywh / (W, H, W, H).  32x32 shared-memory transpose tiles.
__global__ void __launch_bounds__(256) decode_box_transpose_kernel(const float* __restrict__ head, float* __restrict__ out,
                                                                   int no, int anchors, float in_w, float in_h) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int a0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const float* src = head + static_cast<size_t>(b) * no * anchors;
  for (int r = ty; r < 32; r += 8) {
    const int ch = c0 + r, a = a0 + tx;
    if (ch < no && a < anchors) tile[r][tx] = src[static_cast<size_t>(ch) * anchors + a];
  }
  __syncthreads();
  float* dst = out + static_cast<size_t>(b) * anchors * no;
  for (int r = ty; r < 32; r += 8) {
    const int a = a0 + r, ch = c0 + tx;
    if (ch < no && a < anchors) {
      float v = tile[tx][r];
      if (ch < 4) v = v / ((ch & 1) ? in_h : in_w);
      dst[static_cast<size_t>(a) * no + ch] = v;
    }
  }
}

// DecodeBox.decode_box on the upstream 5-tuple (utils/bbox_utils.py:66-82): dbox = the DFL distances (B, 4, A) in grid
// units, cls = class logits (B, nc, A), anchors (2, A), strides (A):
//   dist2bbox(dbox, anchors, xywh, dim 1) * strides  |  sigmoid(cls)  ->  permute(0, 2, 1)  ->  xywh / (W, H, W, H)
// in the reference's float32 operation order (this file is compiled with -fmad=false).  One CTA = 32 anchors of one
// image: channel rows are read along A (coalesced), staged in shared memory and written as 32 contiguous output rows.
__global__ void __launch_bounds__(256) decode_tuple_kernel(const float* __restrict__ dbox, const float* __restrict__ cls,
                                                           const float* __restrict__ anchor_xy, const float* __restrict__ strides,
                                                           float* __restrict__ out, int nc, int anchors, float in_w, float in_h) {
  extern __shared__ float rows[];          // [32][no]
  const int no = 4 + nc;
  const int b = blockIdx.y, a0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 anchors x 8 channel lanes
  const int a = a0 + tx;
  if (a < anchors) {
    if (ty == 0) {
      const float* d = dbox + static_cast<size_t>(b) * 4 * anchors + a;
      const float ax = anchor_xy[a], ay = anchor_xy[anchors + a], st = strides[a];
      const float x1 = __fsub_rn(ax, d[0]), y1 = __fsub_rn(ay, d[anchors]);                       // anchors - lt
      const float x2 = __fadd_rn(ax, d[2 * static_cast<size_t>(anchors)]), y2 = __fadd_rn(ay, d[3 * static_cast<size_t>(anchors)]);
      float* r = rows + tx * no;
      r[0] = __fdiv_rn(__fmul_rn(__fdiv_rn(__fadd_rn(x1, x2), 2.0f), st), in_w);                // ((x1y1 + x2y2) / 2) * s / W
      r[1] = __fdiv_rn(__fmul_rn(__fdiv_rn(__fadd_rn(y1, y2), 2.0f), st), in_h);
      r[2] = __fdiv_rn(__fmul_rn(__fsub_rn(x2, x1), st), in_w);
      r[3] = __fdiv_rn(__fmul_rn(__fsub_rn(y2, y1), st), in_h);
    }
    const float* c = cls + static_cast<size_t>(b) * nc * anchors + a;
    for (int ch = ty; ch < nc; ch += 8) rows[tx * no + 4 + ch] = sigmoid_ref(c[static_cast<size_t>(ch) * anchors]);
  }
  __syncthreads();
  const int n_valid = min(32, anchors - a0) * no;
  float* dst = out + (static_cast<size_t>(b) * anchors + a0) * no;
  for (int i = threadIdx.x; i < n_valid; i += 256) dst[i] = rows[i];
}

// Loss.bbox_decode (model/loss.py:333-337): pred_dist (B, A, 4 * reg_max) logits -> softmax over the reg_max bins of each
// side, projection on arange(reg_max) (`.softmax(3).matmul(proj)`), then dist2bbox(xywh=False) about anchor_points (A, 2):
// (B, A, 4) corner boxes in grid units.  One thread per (image, anchor, side): side 0 / 1 -> anchor - (l, t), side 2 / 3 ->
// anchor + (r, b); a warp reads 8 contiguous anchor rows.  reg_max == 1 (use_dfl False) passes the distances through.
__global__ void __launch_bounds__(256) loss_bbox_decode_kernel(const float* __restrict__ pred_dist, const float* __restrict__ anchor_points,
                                                               float* __restrict__ out, long long total_sides, int anchors, int reg_max) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;      // (b * A + a) * 4 + side
  const bool on = i < total_sides;
  const long long row = (on ? i : 0) >> 2;
  const int side = static_cast<int>(i & 3);
  float dist;
  if (reg_max == 1) {
    dist = on ? pred_dist[i] : 0.f;
  } else {
    const float* l = pred_dist + (on ? i : 0) * reg_max;
    float m = -INFINITY;
    for (int k = 0; k < reg_max; ++k) m = fmaxf(m, l[k]);
    float sum = 0.f, wsum = 0.f;
    for (int k = 0; k < reg_max; ++k) {            // softmax terms e / sum, then the dot with arange: sum_k k * (e_k / sum)
      const float e = expf(__fsub_rn(l[k], m));
      sum = __fadd_rn(sum, e);
    }
    for (int k = 0; k < reg_max; ++k) {
      const float pk = __fdiv_rn(expf(__fsub_rn(l[k], m)), sum);
      wsum = __fadd_rn(wsum, __fmul_rn(pk, static_cast<float>(k)));
    }
    dist = wsum;
  }
  // side 0/1 -> anchor - lt, side 2/3 -> anchor + rb   (dist2bbox, xywh=False, utils/bbox_utils.py:51-55)
  const int a = static_cast<int>(row % anchors);
  const float ap = anchor_points[2 * a + (side & 1)];
  const float v = side < 2 ? __fsub_rn(ap, dist) : __fadd_rn(ap, dist);
  if (on) out[i] = v;
}

}  // namespace tod

using namespace tod;

extern "C" int tod_loss_bbox_decode(const float* d_pred_dist, const float* d_anchor_points, float* d_out, int32_t batch,
                                    int32_t anchors, int32_t reg_max, void* stream) {
  TOD_CHECK_ARG(d_pred_dist && d_anchor_points && d_out, "loss bbox_decode: null pointer");
  TOD_CHECK_ARG(batch > 0 && anchors > 0 && reg_max >= 1 && reg_max <= 64, "loss bbox_decode: bad shape");
  const long long total = static_cast<long long>(batch) * anchors * 4;
  TOD_CHECK_ARG(total < (1ll << 38), "loss bbox_decode: too many boxes");
  loss_bbox_decode_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_pred_dist, d_anchor_points, d_out, total, anchors, reg_max);
  TOD_CHECK_LAUNCH("loss_bbox_decode_kernel launch");
  return TOD_OK;
}

extern "C" int tod_decode_box_from_tuple(const float* d_dbox, const float* d_cls, const float* d_anchors, const float* d_strides,
                                         float* d_decoded, int32_t batch, int32_t nc, int32_t anchors, int32_t in_h,
                                         int32_t in_w, void* stream) {
  TOD_CHECK_ARG(d_dbox && d_cls && d_anchors && d_strides && d_decoded, "decode_box(tuple): null pointer");
  TOD_CHECK_ARG(batch > 0 && batch <= 65535 && nc > 0 && nc <= 4096 && anchors > 0 && in_h > 0 && in_w > 0,
                "decode_box(tuple): bad shape");
  const size_t smem = static_cast<size_t>(32) * (4 + nc) * sizeof(float);
  static PerDeviceOnce attr_once;
  if (attr_once.needed()) {
    int rc = check_cuda(cudaFuncSetAttribute(decode_tuple_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024),
                        "cudaFuncSetAttribute(decode_tuple)");
    if (rc != TOD_OK) return rc;
    attr_once.done();
  }
  TOD_CHECK_ARG(smem <= 200 * 1024, "decode_box(tuple): nc %d too large for shared memory", nc);
  decode_tuple_kernel<<<dim3(ceil_div(anchors, 32), batch), 256, smem, static_cast<cudaStream_t>(stream)>>>(
      d_dbox, d_cls, d_anchors, d_strides, d_decoded, nc, anchors, static_cast<float>(in_w), static_cast<float>(in_h));
  TOD_CHECK_LAUNCH("decode_tuple_kernel launch");
  return TOD_OK;
}

extern "C" int tod_head_decode(const tod_decode_desc* d, void* stream) {
  TOD_CHECK_ARG(d != nullptr, "decode: null descriptor");
  TOD_CHECK_ARG(d->batch > 0 && d->nc > 0 && d->nc <= 1024, "decode: bad batch %d / nc %d", d->batch, d->nc);
  TOD_CHECK_ARG(d->raw_pitch >= 64 + d->nc && d->raw_pitch % 4 == 0, "decode: raw_pitch %d", d->raw_pitch);
  TOD_CHECK_ARG(d->in_h > 0 && d->in_w > 0, "decode: bad input size");
  TOD_CHECK_ARG((d->d_cand_conf == nullptr) == (d->d_cand_cls == nullptr) &&
                    (d->d_cand_conf == nullptr) == (d->d_cand_box == nullptr),
                "decode: candidate outputs must be given together");
  DecodeParams p;
  int total_tiles = 0, anchors = 0;
  for (int l = 0; l < 3; ++l) {
    TOD_CHECK_ARG(d->d_raw[l] != nullptr && d->h[l] > 0 && d->w[l] > 0, "decode: bad level %d", l);
    TOD_CHECK_ARG((reinterpret_cast<uintptr_t>(d->d_raw[l]) & 15) == 0, "decode: raw map %d not 16-byte aligned", l);
    p.raw[l] = d->d_raw[l];
    p.h[l] = d->h[l];
    p.w[l] = d->w[l];
    p.stride[l] = d->stride[l];
    p.tiles[l] = ceil_div(d->h[l] * d->w[l], kDecAnchors);
    p.level_off[l] = anchors;
    anchors += d->h[l] * d->w[l];
    total_tiles += p.tiles[l];
  }
  TOD_CHECK_ARG(!d->d_cand_box || (reinterpret_cast<uintptr_t>(d->d_cand_box) & 15) == 0, "decode: cand_box not aligned");
  TOD_CHECK_ARG(d->batch <= 65535, "decode: batch too large");
  p.raw_pitch = d->raw_pitch;
  p.batch = d->batch;
  p.nc = d->nc;
  p.anchors = anchors;
  p.in_w = static_cast<float>(d->in_w);
  p.in_h = static_cast<float>(d->in_h);
  p.head_out = d->d_head_out;
  p.decoded = d->d_decoded;
  p.cand_box = d->d_cand_box;
  p.cand_conf = d->d_cand_conf;
  p.cand_cls = d->d_cand_cls;
  const bool full_out = d->d_head_out != nullptr || d->d_decoded != nullptr;
  const size_t smem = full_out ? (static_cast<size_t>(kDecAnchors) * (d->nc + 1) + kDecAnchors * 4) * sizeof(float) : 0;
  static PerDeviceOnce attr_once;   // the attribute is per device
  if (attr_once.needed()) {
    int rc = check_cuda(cudaFuncSetAttribute(head_decode_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024),
                        "cudaFuncSetAttribute(head_decode)");
    if (rc != TOD_OK) return rc;
    attr_once.done();
  }
  TOD_CHECK_ARG(smem <= 200 * 1024, "decode: nc %d too large for shared memory", d->nc);
  dim3 grid(total_tiles, d->batch, 1);
  if (full_out)
    head_decode_kernel<true><<<grid, kDecThreads, smem, static_cast<cudaStream_t>(stream)>>>(p);
  else
    head_decode_kernel<false><<<grid, kDecThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
  TOD_CHECK_LAUNCH("head_decode_kernel launch");
  return TOD_OK;
}

extern "C" int tod_decode_box_from_head(const float* d_head_out, float* d_decoded, int32_t batch, int32_t nc,
                                        int32_t anchors, int32_t in_h, int32_t in_w, void* stream) {
  TOD_CHECK_ARG(d_head_out && d_decoded, "decode_box: null pointer");
  TOD_CHECK_ARG(batch > 0 && batch <= 65535 && nc > 0 && anchors > 0 && in_h > 0 && in_w > 0, "decode_box: bad shape");
  const int no = 4 + nc;
  dim3 grid(ceil_div(anchors, 32), ceil_div(no, 32), batch);
  TOD_CHECK_ARG(grid.y <= 65535, "decode_box: too many classes");
  decode_box_transpose_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_head_out, d_decoded, no, anchors, static_cast<float>(in_w), static_cast<float>(in_h));
  TOD_CHECK_LAUNCH("decode_box_transpose_kernel launch");
  return TOD_OK;
}

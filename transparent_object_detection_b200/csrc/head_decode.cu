// Fused head decode: DFL softmax-expectation + anchor/stride box decode + class sigmoid + class max,
// emitting (any subset of) the Head eval tensor, the decode_box tensor and NMS-ready candidates.
// Replaces Head.forward eval branch (model/head.py:53-61), DFL.forward (model/blocks.py:154-157),
// make_anchors (utils/bbox_utils.py:14-37), dist2bbox (:39-58), DecodeBox.decode_box (:66-82) and the
// corner conversion + class max at the top of non_max_suppression (:144-153).
// Arithmetic follows SURVEY.md Appendix B step by step in fp32 (true divisions, no FMA contraction:
// this file is compiled with -fmad=false).  HBM-bound: reads (64+nc) f32 per anchor, writes up to
// 2*(4+nc)+6 f32 per anchor; every global access is a coalesced row segment.
#include "tod_common.cuh"

namespace tod {

constexpr int kDecThreads = 256;
constexpr int kDecAnchors = 64;  // anchors per CTA (4 threads per anchor in the DFL phase)

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

__global__ void __launch_bounds__(kDecThreads) head_decode_kernel(const DecodeParams p) {
  extern __shared__ float dec_smem[];
  const int spitch = p.nc + 1;
  float* s_score = dec_smem;                           // [64][nc + 1]
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

  // ---- phase A: DFL, 4 threads per anchor (one per box side)
  {
    const int a = threadIdx.x >> 2, side = threadIdx.x & 3;
    float dist = 0.f;
    if (a < na) {
      const float4* src = reinterpret_cast<const float4*>(raw + static_cast<size_t>(a) * p.raw_pitch + side * 16);
      float l[16];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 v = __ldg(src + i);
        l[4 * i] = v.x; l[4 * i + 1] = v.y; l[4 * i + 2] = v.z; l[4 * i + 3] = v.w;
      }
      float m = l[0];
#pragma unroll
      for (int i = 1; i < 16; ++i) m = fmaxf(m, l[i]);
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) { l[i] = expf(l[i] - m); s += l[i]; }
#pragma unroll
      for (int i = 0; i < 16; ++i) dist += static_cast<float>(i) * (l[i] / s);   // softmax, then arange(16) projection
    }
    // gather the four sides on the side-0 lane
    const unsigned full = 0xffffffffu;
    const int base = (threadIdx.x & 31) & ~3;
    const float dl = __shfl_sync(full, dist, base + 0);
    const float dt = __shfl_sync(full, dist, base + 1);
    const float dr = __shfl_sync(full, dist, base + 2);
    const float db = __shfl_sync(full, dist, base + 3);
    if (side == 0 && a < na) {
      const int ai = a0 + a;
      const int gy = ai / lw, gx = ai - gy * lw;
      const float ax = static_cast<float>(gx) + 0.5f, ay = static_cast<float>(gy) + 0.5f;  // make_anchors
      const float x1 = ax - dl, y1 = ay - dt, x2 = ax + dr, y2 = ay + db;                 // head.py:57-58
      s_box[a * 4 + 0] = ((x1 + x2) / 2.0f) * stride;                                      // head.py:59-61
      s_box[a * 4 + 1] = ((y1 + y2) / 2.0f) * stride;
      s_box[a * 4 + 2] = (x2 - x1) * stride;
      s_box[a * 4 + 3] = (y2 - y1) * stride;
    }
  }
  // ---- phase B: class sigmoid, coalesced over (anchor, class)
  for (int i = threadIdx.x; i < na * p.nc; i += kDecThreads) {
    const int a = i / p.nc, c = i - a * p.nc;
    const float x = __ldg(raw + static_cast<size_t>(a) * p.raw_pitch + 64 + c);
    s_score[a * spitch + c] = 1.0f / (1.0f + expf(-x));
  }
  __syncthreads();

  const int no = 4 + p.nc;
  const int ag0 = p.level_off[lvl] + a0;  // global anchor index of the CTA's first anchor
  // ---- phase C1: Head eval tensor (B, 4+nc, A): contiguous along anchors
  if (p.head_out) {
    float* o = p.head_out + static_cast<size_t>(b) * no * p.anchors + ag0;
    for (int i = threadIdx.x; i < no * kDecAnchors; i += kDecThreads) {
      const int ch = i / kDecAnchors, a = i - ch * kDecAnchors;
      if (a < na) o[static_cast<size_t>(ch) * p.anchors + a] = ch < 4 ? s_box[a * 4 + ch] : s_score[a * spitch + ch - 4];
    }
  }
  // ---- phase C2: decode_box tensor (B, A, 4+nc): contiguous rows, xywh / (W,H,W,H)
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
  // ---- phase C3: NMS candidates: corners of the normalised xywh (bbox_utils.py:144-148), class max (:153)
  if (p.cand_conf && threadIdx.x < na) {
    const int a = threadIdx.x;
    float best = s_score[a * spitch];
    int bi = 0;
    for (int c = 1; c < p.nc; ++c) {
      const float v = s_score[a * spitch + c];
      if (v > best) { best = v; bi = c; }   // strict: first (lowest) class wins ties, like torch.max on CPU
    }
    const size_t g = static_cast<size_t>(b) * p.anchors + ag0 + a;
    p.cand_conf[g] = best;
    p.cand_cls[g] = bi;
    const float nx = s_box[a * 4 + 0] / p.in_w, ny = s_box[a * 4 + 1] / p.in_h;
    const float nw = s_box[a * 4 + 2] / p.in_w, nh = s_box[a * 4 + 3] / p.in_h;
    reinterpret_cast<float4*>(p.cand_box)[g] = make_float4(nx - nw / 2.0f, ny - nh / 2.0f, nx + nw / 2.0f, ny + nh / 2.0f);
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

}  // namespace tod

using namespace tod;

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
  const size_t smem = (static_cast<size_t>(kDecAnchors) * (d->nc + 1) + kDecAnchors * 4) * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    int rc = check_cuda(cudaFuncSetAttribute(head_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024),
                        "cudaFuncSetAttribute(head_decode)");
    if (rc != TOD_OK) return rc;
    attr_done = true;
  }
  TOD_CHECK_ARG(smem <= 200 * 1024, "decode: nc %d too large for shared memory", d->nc);
  dim3 grid(total_tiles, d->batch, 1);
  head_decode_kernel<<<grid, kDecThreads, smem, static_cast<cudaStream_t>(stream)>>>(p);
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

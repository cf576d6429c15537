// Arithmetic of the head decode, shared by head_decode.cu (raw maps -> tensors / candidates) and by the conv kernel's
// fused head-output epilogues (conv_halo_tcgen05.cu), so that both paths produce bit-identical candidates.
// Every floating-point step is an explicit round-to-nearest intrinsic: no FMA contraction whatever the compile flags.
// Follows SURVEY.md Appendix B: DFL.forward model/blocks.py:154-157, make_anchors utils/bbox_utils.py:14-37,
// Head.forward eval branch model/head.py:53-61, non_max_suppression prologue utils/bbox_utils.py:144-153.
#pragma once

#include "tod_common.cuh"

namespace tod {

__device__ __forceinline__ float sigmoid_ref(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }   // head.py:61

// (score, class) candidates are ordered by score descending, then class ascending (torch.max keeps the first maximum)
__device__ __forceinline__ void better(float& s, int& c, float s2, int c2) {
  if (s2 > s || (s2 == s && c2 < c)) { s = s2; c = c2; }
}

// softmax over the 16 bins of one box side, then the arange(16) projection (blocks.py:155-157).
// exp is the ex2-based __expf (2 ulp for the near-maximum bins that carry the weight, like expf; only bins with negligible
// weight lose more) and the 16 per-bin divisions of softmax are one IEEE reciprocal and 16 multiplies (<= 1 ulp per
// term): 4x fewer instructions, |delta distance| ~ 1e-6 bins, far inside the decode tolerance (tests: 2e-3 px).
__device__ __forceinline__ float dfl_side(float (&l)[16]) {
  float m = fmaxf(fmaxf(fmaxf(l[0], l[1]), fmaxf(l[2], l[3])), fmaxf(fmaxf(l[4], l[5]), fmaxf(l[6], l[7])));
  m = fmaxf(m, fmaxf(fmaxf(fmaxf(l[8], l[9]), fmaxf(l[10], l[11])), fmaxf(fmaxf(l[12], l[13]), fmaxf(l[14], l[15]))));
#pragma unroll
  for (int i = 0; i < 16; ++i) l[i] = __expf(__fsub_rn(l[i], m));
  // pairwise sums (independent chains); every step is an explicit rn add / mul, identical in every translation unit
  float s4[4], w4[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = 4 * q;
    s4[q] = __fadd_rn(__fadd_rn(l[i], l[i + 1]), __fadd_rn(l[i + 2], l[i + 3]));
    w4[q] = __fadd_rn(__fadd_rn(__fmul_rn(static_cast<float>(i), l[i]), __fmul_rn(static_cast<float>(i + 1), l[i + 1])),
                      __fadd_rn(__fmul_rn(static_cast<float>(i + 2), l[i + 2]), __fmul_rn(static_cast<float>(i + 3), l[i + 3])));
  }
  const float sum = __fadd_rn(__fadd_rn(s4[0], s4[1]), __fadd_rn(s4[2], s4[3]));
  const float wsum = __fadd_rn(__fadd_rn(w4[0], w4[1]), __fadd_rn(w4[2], w4[3]));
  return __fmul_rn(wsum, __fdiv_rn(1.0f, sum));
}

// ltrb distances about anchor (gx + 0.5, gy + 0.5) -> xywh in input pixels (head.py:57-61)
__device__ __forceinline__ float4 box_xywh_px(float dl, float dt, float dr, float db, int gx, int gy, float stride) {
  const float ax = __fadd_rn(static_cast<float>(gx), 0.5f), ay = __fadd_rn(static_cast<float>(gy), 0.5f);   // make_anchors
  const float x1 = __fsub_rn(ax, dl), y1 = __fsub_rn(ay, dt), x2 = __fadd_rn(ax, dr), y2 = __fadd_rn(ay, db);
  return make_float4(__fmul_rn(__fdiv_rn(__fadd_rn(x1, x2), 2.0f), stride), __fmul_rn(__fdiv_rn(__fadd_rn(y1, y2), 2.0f), stride),
                     __fmul_rn(__fsub_rn(x2, x1), stride), __fmul_rn(__fsub_rn(y2, y1), stride));
}

// NMS candidate corners of the normalised xywh (bbox_utils.py:77-82 then :144-148)
__device__ __forceinline__ float4 box_corners_norm(const float4 b, float in_w, float in_h) {
  const float nx = __fdiv_rn(b.x, in_w), ny = __fdiv_rn(b.y, in_h), nw = __fdiv_rn(b.z, in_w), nh = __fdiv_rn(b.w, in_h);
  const float hw = __fdiv_rn(nw, 2.0f), hh = __fdiv_rn(nh, 2.0f);
  return make_float4(__fsub_rn(nx, hw), __fsub_rn(ny, hh), __fadd_rn(nx, hw), __fadd_rn(ny, hh));
}

// Class maximum found on the logits: the two largest logits seen (first occurrence wins ties, like an ascending-class
// scan of the scores) ...
struct LogitTop2 {
  float x1 = -INFINITY, x2 = -INFINITY;
  int c1 = 0x7fffffff;
  __device__ __forceinline__ void add(float x, int c) {          // branch-free: 1 compare, 1 max, 3 selects
    const bool gt = x > x1;
    x2 = gt ? x1 : fmaxf(x2, x);
    c1 = gt ? c : c1;
    x1 = gt ? x : x1;
  }
  // merge a tracker that saw a disjoint set of classes: largest value wins, lowest class among equal values;
  // x2 = second largest of the union (duplicates of the maximum count)
  __device__ __forceinline__ void merge(const LogitTop2& o) {
    const bool take = o.x1 > x1 || (o.x1 == x1 && o.c1 < c1);
    const float lose = take ? x1 : o.x1;
    x2 = fmaxf(fmaxf(x2, o.x2), lose);
    c1 = take ? o.c1 : c1;
    x1 = take ? o.x1 : x1;
  }
};
// ... and the logit threshold below which a class cannot tie (or, through rounding, overtake) the best float32 score.
// d sigmoid / dx = s (1 - s): a logit gap dx changes the score by dx (1 - s) relative, i.e. dx (1 - s) / 6e-8 ulps, against
// <= ~4 ulps of evaluation error per score (expf 2 ulp, add and divide 0.5 ulp each).  m <= 2 (1 - s >= 0.12): 1e-4 is
// 200 ulps; m <= 8 (1 - s >= 3.3e-4): 0.01 is 55 ulps; m < 15 (1 - s >= 3e-7): 2.0; outside (-80, 15) the sigmoid
// saturates and every class counts.  (The window decides only WHICH anchors take the exact rescan; with 0.01 everywhere
// ~3 % of random-logit anchors, i.e. 60 % of the warps, took it.)
__device__ __forceinline__ float tie_window_threshold(float m) {
  return (m > -80.0f && m < 15.0f) ? __fsub_rn(m, m > 8.0f ? 2.0f : (m > 2.0f ? 0.01f : 1e-4f)) : -INFINITY;
}

}  // namespace tod

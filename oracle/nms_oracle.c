/*
 * CPU oracle (plain C) of the detector's post-processing arithmetic -- TEST INFRASTRUCTURE, NOT THE PRODUCT.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * Restates, in the reference's float32 operation order:
 *   - torchvision.ops.nms (CPU kernel semantics), called by DecodeBox.non_max_suppression at
 *     /root/reference/utils/bbox_utils.py:172 -- stable descending sort, greedy suppression when the float32 IoU,
 *     widened to double, is > the double threshold; zero-area pairs give NaN and survive;
 *   - the reference's per-image / per-class loop around it, utils/bbox_utils.py:144-175 (xywh -> corners in
 *     float32 as cx - w/2 ..., class max with lowest-id tie break, conf >= thres in float32, classes ascending).
 * Pinned by tests/test_oracle_golden.py against fixtures generated from the reference itself
 * (oracle/make_golden.py, torchvision 0.26.0+cu128).
 *
 * build: make -C oracle   (gcc -O2 -fno-fast-math -ffp-contract=off: no FMA contraction, no reassociation)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* stable descending order of scores[0..n) -> order[] */
static void stable_argsort_desc(const float* scores, int32_t n, int32_t* order, int32_t* tmp) {
  for (int32_t i = 0; i < n; ++i) order[i] = i;
  for (int32_t width = 1; width < n; width *= 2) { /* bottom-up merge sort: stable */
    for (int32_t lo = 0; lo < n; lo += 2 * width) {
      int32_t mid = lo + width < n ? lo + width : n, hi = lo + 2 * width < n ? lo + 2 * width : n;
      int32_t a = lo, b = mid, k = lo;
      while (a < mid && b < hi) tmp[k++] = (scores[order[b]] > scores[order[a]]) ? order[b++] : order[a++];
      while (a < mid) tmp[k++] = order[a++];
      while (b < hi) tmp[k++] = order[b++];
    }
    memcpy(order, tmp, (size_t)n * sizeof(int32_t));
  }
}

/* torchvision.ops.nms: boxes [n][4] (x1,y1,x2,y2) float32, scores [n]; writes kept indices (score-descending) to
 * keep[], returns their count. */
int32_t tod_oracle_nms(const float* boxes, const float* scores, int32_t n, double iou_thr, int32_t* keep) {
  if (n <= 0) return 0;
  int32_t* order = (int32_t*)malloc((size_t)n * sizeof(int32_t));
  int32_t* tmp = (int32_t*)malloc((size_t)n * sizeof(int32_t));
  float* areas = (float*)malloc((size_t)n * sizeof(float));
  uint8_t* suppressed = (uint8_t*)calloc((size_t)n, 1);
  stable_argsort_desc(scores, n, order, tmp);
  for (int32_t i = 0; i < n; ++i) {
    const float* b = boxes + 4 * (size_t)i;
    areas[i] = (b[2] - b[0]) * (b[3] - b[1]);
  }
  int32_t nk = 0;
  for (int32_t _i = 0; _i < n; ++_i) {
    const int32_t i = order[_i];
    if (suppressed[i]) continue;
    keep[nk++] = i;
    const float ix1 = boxes[4 * (size_t)i], iy1 = boxes[4 * (size_t)i + 1], ix2 = boxes[4 * (size_t)i + 2],
                iy2 = boxes[4 * (size_t)i + 3], iarea = areas[i];
    for (int32_t _j = _i + 1; _j < n; ++_j) {
      const int32_t j = order[_j];
      if (suppressed[j]) continue;
      const float* bj = boxes + 4 * (size_t)j;
      const float xx1 = ix1 > bj[0] ? ix1 : bj[0];
      const float yy1 = iy1 > bj[1] ? iy1 : bj[1];
      const float xx2 = ix2 < bj[2] ? ix2 : bj[2];
      const float yy2 = iy2 < bj[3] ? iy2 : bj[3];
      float w = xx2 - xx1, h = yy2 - yy1;
      w = w > 0.0f ? w : 0.0f;
      h = h > 0.0f ? h : 0.0f;
      const float inter = w * h;
      const float ovr = inter / (iarea + areas[j] - inter);
      if ((double)ovr > iou_thr) suppressed[j] = 1; /* NaN compares false: zero-area pairs survive */
    }
  }
  free(order);
  free(tmp);
  free(areas);
  free(suppressed);
  return nk;
}

/* One image of DecodeBox.non_max_suppression's selection (utils/bbox_utils.py:144-175): prediction [anchors][4+nc]
 * normalised xywh + class scores (not modified); writes kept anchor indices in the reference's output order (class
 * ascending, score descending within a class) to keep[], returns their count. */
int32_t tod_oracle_nms_image(const float* prediction, int32_t anchors, int32_t nc, float conf_thres, double iou_thr,
                             int32_t* keep) {
  const int32_t no = 4 + nc;
  float* corner = (float*)malloc((size_t)anchors * 4 * sizeof(float));
  float* conf = (float*)malloc((size_t)anchors * sizeof(float));
  int32_t* cls = (int32_t*)malloc((size_t)anchors * sizeof(int32_t));
  int32_t* seg = (int32_t*)malloc((size_t)anchors * sizeof(int32_t));
  float* seg_box = (float*)malloc((size_t)anchors * 4 * sizeof(float));
  float* seg_conf = (float*)malloc((size_t)anchors * sizeof(float));
  int32_t* seg_keep = (int32_t*)malloc((size_t)anchors * sizeof(int32_t));
  for (int32_t a = 0; a < anchors; ++a) {
    const float* p = prediction + (size_t)a * no;
    corner[4 * a + 0] = p[0] - p[2] / 2.0f;
    corner[4 * a + 1] = p[1] - p[3] / 2.0f;
    corner[4 * a + 2] = p[0] + p[2] / 2.0f;
    corner[4 * a + 3] = p[1] + p[3] / 2.0f;
    int32_t best = 0;
    for (int32_t c = 1; c < nc; ++c)
      if (p[4 + c] > p[4 + best]) best = c; /* first maximum: lowest class id on ties */
    cls[a] = best;
    conf[a] = p[4 + best];
  }
  int32_t nk = 0;
  for (int32_t c = 0; c < nc; ++c) {
    int32_t ns = 0;
    for (int32_t a = 0; a < anchors; ++a)
      if (cls[a] == c && conf[a] >= conf_thres) {
        seg[ns] = a;
        memcpy(seg_box + 4 * (size_t)ns, corner + 4 * (size_t)a, 4 * sizeof(float));
        seg_conf[ns] = conf[a];
        ++ns;
      }
    const int32_t k = tod_oracle_nms(seg_box, seg_conf, ns, iou_thr, seg_keep);
    for (int32_t i = 0; i < k; ++i) keep[nk++] = seg[seg_keep[i]];
  }
  free(corner);
  free(conf);
  free(cls);
  free(seg);
  free(seg_box);
  free(seg_conf);
  free(seg_keep);
  return nk;
}

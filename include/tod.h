/*
 * tod.h -- C ABI of libtod.so: the B200 (sm_100a) detector-inference hot path.
 *
 * The reference (mohamed22311/Transparent-Object-Detection) is pure Python/PyTorch and has no
 * FFI layer: the boundary a maintainer binds is the set of plain-C entry
 * points below, called through ctypes from Python classes that keep the reference's own call
 * surface (BaseModel / DecodeBox).  Each entry point cites the reference
 * code it replaces.
 *
 * Conventions
 *   - all pointers named d_* are DEVICE pointers owned by the caller (PyTorch's allocator);
 *     the library never allocates, frees or synchronises;
 *   - kernels are enqueued on `stream` (a cudaStream_t passed as void*);
 *   - return value: 0 = ok, negative = error; tod_last_error() gives a thread-local message;
 *   - no C++ exceptions cross the ABI; there is no CPU fallback of any kind.
 *   - activations are NHWC bf16 ("pixels x channels", channel pitch may exceed the logical
 *     channel count so that producers write straight into concat buffers at a channel offset).
 */
#ifndef TOD_H_
#define TOD_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TOD_OK 0
#define TOD_ERR_INVALID (-1)
#define TOD_ERR_CUDA (-2)
#define TOD_ERR_UNSUPPORTED (-3)

#define TOD_ACT_NONE 0
#define TOD_ACT_SILU 1

#define TOD_OUT_BF16 0
#define TOD_OUT_F32 1

/* library / build identification */
int tod_version(void);
const char* tod_last_error(void);
/* 1 when the current device is compute capability 10.x, else 0 (negative on CUDA error). */
int tod_device_ok(void);

/*
 * Fused Conv2d(bias=False) + BatchNorm2d(eval) + SiLU [+ residual] as an NHWC bf16 implicit GEMM on
 * tcgen05/TMEM fed by TMA.
 * Replaces: Conv.forward  model/blocks.py:52-54 (with fuse_conv :160-187 folded into w/bias),
 *           Bottleneck.forward's shortcut add  model/blocks.py:80-82 (residual),
 *           the bare nn.Conv2d 1x1 + bias of the head  model/head.py:31,42 (act = NONE, out = F32),
 *           torch.cat / chunk in C2f/SPPF/Neck  model/blocks.py:106-108,139-142, model/neck.py:57-60
 *           (x/out/residual are channel-offset views of wider buffers),
 *           nn.Upsample(2,'nearest') feeding a 1x1 conv  model/neck.py:17,57-58 (upadd: the 1x1 conv of
 *           the low-resolution operand is computed at low resolution and added pre-activation at
 *           [h>>1][w>>1]).
 * GEMM view: M = batch*Hout*Wout pixels, N = cout, K = ksize*ksize*cin.
 * Supported: ksize 1 (stride 1) or 3 (stride 1 or 2, pad 1); cin % 16 == 0; cout % 16 == 0; pitches % 8 == 0;
 *            stride 2 needs even hin/win.
 */
typedef struct tod_conv_desc {
  const void* d_x;        /* bf16 [batch, hin, win, x_pitch] (already offset to the first input channel) */
  const void* d_w;        /* bf16 [cout, ksize*ksize*cin_pad], K index = tap*cin_pad + c  (tod_conv_weight_layout) */
  const float* d_bias;    /* f32 [cout] or NULL */
  const void* d_residual; /* bf16 [batch, hout, wout, res_pitch] added AFTER activation, or NULL */
  const float* d_upadd;   /* f32 [batch, hout/2, wout/2, cout] added BEFORE activation at [h>>1][w>>1], or NULL */
  void* d_out;            /* bf16 or f32 [batch, hout, wout, out_pitch] (already offset to the first output channel) */
  int32_t batch, hin, win, cin, cout;
  int32_t ksize, stride;
  int32_t x_pitch, res_pitch, out_pitch; /* in elements */
  int32_t act;            /* TOD_ACT_* */
  int32_t out_dtype;      /* TOD_OUT_* */
  int32_t block_k;        /* 0 = auto (64/32/16 from cin) */
  int32_t num_stages;     /* 0 = auto */
  int32_t reserved[4];    /* tools only: kernel variant, m, no resident weights, N-tile cap (tools/conv_bench.py) */
  int32_t flags;          /* TOD_CONV_* bits below */
  int32_t reserved2[3];
} tod_conv_desc;

/* d_w was written by an earlier kernel of the same stream (a GEMM whose "weights" are activations, e.g. q . k^T of
 * SelfAttention, model/blocks.py:243-245): the kernel must not prefetch it ahead of the programmatic-launch wait. */
#define TOD_CONV_DYNAMIC_W 1
/* Walk the output tiles from the last image to the first.  Consecutive layers of the plan alternate direction so that a
 * layer starts on the activations its producer wrote last, which are the ones still resident in the 126 MB L2. */
#define TOD_CONV_REVERSE 2
/* Tests / tools: keep the 16x8 pixel patches where the plan would pick row-flat tiles (both give bit-identical outputs). */
#define TOD_CONV_PATCH_TILES 4
/* Tests / tools: force the CTA-pair kernel (tcgen05 cta_group::2: two SMs share every weight tile) on / off where the
 * plan would decide by itself.  Both give bit-identical outputs (same products, same accumulation order). */
#define TOD_CONV_PAIR_ON 8
#define TOD_CONV_PAIR_OFF 16

/*
 * SM budget of the persistent kernels' grids (convs, stem): 0 = the whole device (default; TOD_SM_BUDGET overrides the
 * default).  Read at enqueue / graph-capture time.  K batches in flight on K streams with a budget of #SMs / K each run side
 * by side on disjoint SMs, every CTA living K times longer: the per-launch cost of a CTA (prologue, wait for the previous
 * grid, first-load latency, drain of the last accumulator) is amortised over K times more tiles.  The reference has no
 * counterpart (its layers are cuDNN / ATen launches, model/blocks.py:52-54); results do not depend on it.
 */
int tod_set_sm_budget(int32_t sms);
int tod_get_sm_budget(void);

int tod_conv2d_nhwc_bf16(const tod_conv_desc* desc, void* stream);

/*
 * The last conv of a head tower fused with the head decode (detect path: only NMS candidates are needed).
 * Replaces, in one launch and without materialising the raw (64 + nc)-channel maps: the bare 1x1 conv + bias
 * model/head.py:31 (box tower) or :42 (class tower) together with its share of Head.forward's eval branch
 * model/head.py:53-61, DFL.forward model/blocks.py:154-157, make_anchors utils/bbox_utils.py:14-37,
 * DecodeBox.decode_box :77-82 and the corner conversion / class max of non_max_suppression :144-153.
 *   mode TOD_FUSE_BOX: conv cout must be 64 (4 x 16 DFL bins)  -> d_cand_box f32 [batch, A, 4]
 *   mode TOD_FUSE_CLS: conv cout >= nc (padded to 16)          -> d_cand_conf f32 [batch, A], d_cand_cls i32 [batch, A]
 * desc must be a 1x1, stride 1, act NONE conv (d_out is ignored and may be NULL); level_off is the index of the
 * level's first anchor among the A anchors of an image, stride = in_h / hin.  The candidates are bit-identical to
 * tod_conv2d_nhwc_bf16 (f32 out) followed by tod_head_decode.
 */
enum { TOD_FUSE_BOX = 1, TOD_FUSE_CLS = 2 };
typedef struct tod_head_fuse_desc {
  int32_t mode;
  int32_t nc;
  int32_t level_off, anchors;
  int32_t in_h, in_w;
  float stride;
  float* d_cand_box;
  float* d_cand_conf;
  int32_t* d_cand_cls;
  int32_t reserved[4];
} tod_head_fuse_desc;

int tod_conv2d_head_decode(const tod_conv_desc* desc, const tod_head_fuse_desc* fuse, void* stream);

/*
 * A conv followed by a 1x1 conv on its output, the intermediate kept on chip ("back-to-back GEMM"): the first conv's
 * activated bf16 tile is written to shared memory in the tcgen05 A-operand layout and a second MMA with the resident
 * 1x1 weights produces the final tile; the intermediate tensor is never written to or read from HBM.
 * Replaces: a Conv (model/blocks.py:52-54) whose only consumer is C2f.cv1 (model/blocks.py:104), e.g.
 *           backbone.dark2[0] -> backbone.dark2[1].cv1 (model/backbone.py:23-25), bit-identical to the two separate calls.
 * Supported: first conv with 64 output channels, bf16 out, SiLU, no residual / upadd; tail 64 -> 64, SiLU.
 *   desc: the first conv (d_out ignored);  d_w2 bf16 [64, 64] (tod_conv_weight_layout for cin 64, ksize 1), d_bias2 f32 [64]
 *   d_out2 bf16 [batch, hout, wout, out2_pitch]
 */
typedef struct tod_conv_tail_desc {
  const void* d_w2;
  const float* d_bias2;
  void* d_out2;
  int32_t cout2, out2_pitch, act2;
  int32_t reserved[5];
} tod_conv_tail_desc;
int tod_conv2d_tail1x1(const tod_conv_desc* desc, const tod_conv_tail_desc* tail, void* stream);
/* The same with the tail being the bare 64 -> 64 box-logit conv of a head tower (model/head.py:36-42, act2 = NONE) and its
 * DFL / dist2bbox decode (tod_conv2d_head_decode, mode TOD_FUSE_BOX): Conv3x3 + Conv1x1 + decode in one kernel, neither
 * the tower's last activation nor the logits reach memory; candidates bit-identical to the separate calls. */
int tod_conv2d_tail1x1_box_decode(const tod_conv_desc* desc, const tod_conv_tail_desc* tail,
                                  const tod_head_fuse_desc* fuse, void* stream);

/* Packed-weight geometry for a conv: *block_k (TMA/UMMA K chunk), *cin_pad (cin rounded up to block_k),
 * *k_total = ksize*ksize*cin_pad.  Host packers lay weights out as [cout][tap][cin_pad] bf16, zero padded. */
int tod_conv_weight_layout(int32_t cin, int32_t ksize, int32_t block_k_hint,
                           int32_t* block_k, int32_t* cin_pad, int32_t* k_total);

/*
 * Stem: Conv2d(3, cout, 3, stride 2, pad 1) + BN + SiLU reading the caller's NCHW fp32 image tensor and
 * writing NHWC bf16.  Replaces backbone.stem  model/backbone.py:20 (Conv, model/blocks.py:52-54) and the
 * layout/dtype conversion in front of it.
 *   d_x   f32 [batch, 3, hin, win] (device)
 *   h_w   f32 [cout, 27] HOST pointer (BN folded, K index = c*9 + kh*3 + kw), h_bias f32 [cout] HOST pointer or NULL:
 *         the 3.5..14 KB of folded weights travel as a kernel parameter (constant bank); cout in {16,32,48,64,96,128}
 *   d_out bf16 [batch, hin/2, win/2, out_pitch] (device)
 */
int tod_stem_conv_nchw_f32(const float* d_x, const float* h_w, const float* h_bias, void* d_out,
                           int32_t batch, int32_t hin, int32_t win, int32_t cout, int32_t out_pitch,
                           void* stream);

/*
 * Stem from uint8 images on the tensor cores, with the reference's pre-processing fused in: replaces
 * `np.array(image, float32) / 255.0` + HWC->CHW (utils/callbacks.py:142-144, dataset/coco/get_map.py:52-60) followed by
 * backbone.stem (model/backbone.py:20).  Pixel values are exact in bf16; 1/255 is folded into the bf16 weights.
 *   d_x   u8 [batch, hin, win, 3] (device, the letterboxed RGB images)
 *   h_w, h_bias  as for tod_stem_conv_nchw_f32 (HOST pointers, BN folded, K index = c*9 + kh*3 + kw)
 *   d_out bf16 [batch, hin/2, win/2, out_pitch] (device)
 */
int tod_stem_conv_nhwc_u8(const uint8_t* d_x, const float* h_w, const float* h_bias, void* d_out,
                          int32_t batch, int32_t hin, int32_t win, int32_t cout, int32_t out_pitch,
                          void* stream);

/*
 * SPPF pooling: three chained MaxPool2d(5, 1, 2) evaluated in shared memory.
 * Replaces SPPF.forward's pooling + torch.cat  model/blocks.py:139-141.
 * d_buf is the bf16 concat buffer [batch, h, w, pitch] whose channels [0, c) already hold cv1's output;
 * channels [c,2c), [2c,3c), [3c,4c) receive the 5x5, 9x9, 13x13 max-pools.
 */
int tod_sppf_pool_nhwc_bf16(void* d_buf, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t pitch,
                            void* stream);

/*
 * Head decode: DFL softmax-expectation, ltrb->xywh about the anchor, x stride, class sigmoid, class max.
 * Replaces Head.forward eval branch  model/head.py:53-61, DFL.forward  model/blocks.py:154-157,
 * make_anchors  utils/bbox_utils.py:14-37, dist2bbox :39-58 and DecodeBox.decode_box :66-82.
 * Inputs: per level l the raw map d_raw[l] = f32 [batch, h_l*w_l, raw_pitch] with channels [0,64) box bins and
 * [64, 64+nc) class logits (what Head returns in training mode, NHWC).
 * Outputs (any may be NULL):
 *   d_head_out f32 [batch, 4+nc, A]   -- Head eval tensor: xywh in input pixels, sigmoid scores
 *   d_decoded  f32 [batch, A, 4+nc]   -- decode_box: xywh / (W,H,W,H), scores
 *   d_cand_box f32 [batch, A, 4]      -- NMS-ready corners (cx-w/2, cy-h/2, cx+w/2, cy+h/2 of the normalised xywh,
 *                                        utils/bbox_utils.py:144-148)
 *   d_cand_conf f32 [batch, A], d_cand_cls i32 [batch, A] -- class max (lowest class id on ties) and its index
 */
typedef struct tod_decode_desc {
  const float* d_raw[3];
  int32_t h[3], w[3];
  float stride[3];
  int32_t raw_pitch;
  int32_t batch, nc;
  int32_t in_h, in_w;  /* network input size (normalisation) */
  float* d_head_out;
  float* d_decoded;
  float* d_cand_box;
  float* d_cand_conf;
  int32_t* d_cand_cls;
  int32_t reserved[4];
} tod_decode_desc;

int tod_head_decode(const tod_decode_desc* desc, void* stream);

/*
 * DecodeBox.decode_box applied to an existing Head eval tensor (utils/bbox_utils.py:77-82):
 * d_head_out f32 [batch, 4+nc, A] -> d_decoded f32 [batch, A, 4+nc] with xywh / (in_w, in_h, in_w, in_h).
 */
int tod_decode_box_from_head(const float* d_head_out, float* d_decoded, int32_t batch, int32_t nc, int32_t anchors,
                             int32_t in_h, int32_t in_w, void* stream);

/*
 * Loss.bbox_decode (model/loss.py:333-337; SURVEY 8 row f4, forward only): pred_dist f32 [batch, anchors, 4 * reg_max] logits
 * (the box half of the training-mode head maps, model/head.py:50-51, flattened and permuted as loss.py:343-347 does) ->
 * softmax over the reg_max bins of each side . arange(reg_max) -> dist2bbox(xywh=False) about anchor_points f32 [anchors, 2]
 * (make_anchors, utils/bbox_utils.py:14-37): out f32 [batch, anchors, 4] = (x1, y1, x2, y2) in grid units.
 * reg_max == 1: the distances are taken as they are (use_dfl False).
 */
int tod_loss_bbox_decode(const float* d_pred_dist, const float* d_anchor_points, float* d_out, int32_t batch, int32_t anchors,
                         int32_t reg_max, void* stream);

/*
 * DecodeBox.decode_box on the upstream 5-tuple the reference's callers are written for (utils/bbox_utils.py:66-82;
 * callers utils/callbacks.py:150-151, dataset/coco/get_map.py:68-69): (dbox, cls, origin_cls, anchors, strides) ->
 * d_decoded f32 [batch, A, 4+nc] = cat(dist2bbox(dbox, anchors, xywh) * strides, sigmoid(cls)).permute(0, 2, 1) with
 * xywh / (in_w, in_h, in_w, in_h).  Box columns are bit-exact with the reference (same float32 operation order);
 * origin_cls is unused by the reference as well.
 *   d_dbox f32 [batch, 4, A] (DFL distances l, t, r, b in grid units), d_cls f32 [batch, nc, A] (logits),
 *   d_anchors f32 [2, A] (x row, y row: make_anchors(...)[0].transpose(0, 1), model/head.py:53), d_strides f32 [A]
 */
int tod_decode_box_from_tuple(const float* d_dbox, const float* d_cls, const float* d_anchors, const float* d_strides,
                              float* d_decoded, int32_t batch, int32_t nc, int32_t anchors, int32_t in_h, int32_t in_w,
                              void* stream);

/*
 * NMS, step 0 (only for the dense reference-layout tensor): xywh -> corners IN PLACE on
 * d_prediction f32 [batch, A, 4+nc] (the reference's side effect, utils/bbox_utils.py:144-149) and class max.
 * Replaces utils/bbox_utils.py:144-153.
 */
int tod_nms_prepare_dense(float* d_prediction, int32_t batch, int32_t anchors, int32_t nc,
                          float* d_cand_box, float* d_cand_conf, int32_t* d_cand_cls, void* stream);

/*
 * NMS proper: confidence filter (conf >= conf_thres, float32 compare), per-class greedy suppression with
 * torchvision.ops.nms semantics (IoU in float32, suppressed when (double)IoU > iou_thres), output ordered
 * class-ascending then score-descending (ties: lower anchor index first).
 * Replaces DecodeBox.non_max_suppression's per-image / per-class loops  utils/bbox_utils.py:151-175 and
 * torchvision.ops.nms (:172).
 *   inputs   d_cand_box f32 [batch, A, 4], d_cand_conf f32 [batch, A], d_cand_cls i32 [batch, A]
 *   workspace d_work: tod_nms_workspace_bytes(batch, A) bytes
 *   outputs  d_keep_idx i32 [batch, A] (anchor indices in output order; first d_keep_count[b] valid)
 *            d_keep_count i32 [batch]
 *            d_dets f32 [batch, A, 6] rows [x1, y1, x2, y2, conf, cls] in output order, or NULL
 */
int64_t tod_nms_workspace_bytes(int32_t batch, int32_t anchors);
int tod_nms(const float* d_cand_box, const float* d_cand_conf, const int32_t* d_cand_cls,
            int32_t batch, int32_t anchors, float conf_thres, double iou_thres,
            void* d_work, int64_t work_bytes,
            int32_t* d_keep_idx, int32_t* d_keep_count, float* d_dets, void* stream);

/*
 * Un-letterbox the kept rows on the device.
 * Replaces: the tail of non_max_suppression  utils/bbox_utils.py:176-180 + correct_boxes :84-117 (numpy on the host in
 *           the reference), same dtype flow (float32 centres/sizes, float64 letterbox arithmetic, float32 rows).
 *   d_dets f32 [batch, anchors, 6] rows [x1, y1, x2, y2, conf, cls] (tod_nms), first d_keep_count[b] valid per image
 *   d_params f64 [batch, 6] = offset_y, offset_x, scale_y, scale_x, image_h, image_w per image, where
 *            new = round(image * min(input / image)), offset = (input - new) / 2 / input, scale = input / new (:105-107)
 *            (offset 0, scale 1 when letterbox == 0)
 *   d_rows  f32 [batch, anchors, 6] rows [y1, x1, y2, x2, conf, cls] in image pixels (may alias d_dets)
 */
int tod_correct_boxes(const float* d_dets, const int32_t* d_keep_count, int32_t batch, int32_t anchors,
                      const double* d_params, int32_t letterbox, float* d_rows, void* stream);

/*
 * Result packing + top-`max_boxes` on the device: the batch's kept rows compacted image after
 * image, so that one small device-to-host copy carries every detection.
 * Replaces: the per-image `.cpu().numpy()` of utils/bbox_utils.py:178 and, with max_boxes > 0, the top-k of the reference's
 *           detect loop, utils/callbacks.py:159-166 (`np.argsort(top_conf)[::-1][:self.max_boxes]`; dataset/coco/get_map.py:81-94
 *           keeps every row = max_boxes 0).
 *   d_rows f32 [batch, anchors, 6] (tod_nms dets or tod_correct_boxes rows), first d_keep_count[b] valid per image
 *   max_boxes <= 0: rows keep the NMS output order;  > 0: at most max_boxes rows per image, score descending, ties in kept
 *                   order (numpy's default argsort is unstable, so the reference defines no tie order)
 *   d_offsets i32 [batch + 1]: image b's rows are d_packed[d_offsets[b] .. d_offsets[b + 1]); d_offsets[batch] = total
 *   d_packed f32 [sum, 6] (capacity batch * anchors rows, or batch * max_boxes)
 *   d_work: tod_pack_workspace_bytes(batch, anchors) bytes (0 unless anchors > 16384 and max_boxes > 0)
 */
int64_t tod_pack_workspace_bytes(int32_t batch, int32_t anchors);
int tod_pack_detections(const float* d_rows, const int32_t* d_keep_count, int32_t batch, int32_t anchors, int32_t max_boxes,
                        int32_t* d_offsets, float* d_packed, void* d_work, int64_t work_bytes, void* stream);

/*
 * Letterbox preprocessing on the device.
 * Replaces: resize_image  utils/utils.py:16-30 (Pillow Image.resize(size, Image.BICUBIC) + Image.new((128,128,128)) +
 *           paste at ((w-nw)//2, (h-nh)//2)) as called by utils/callbacks.py:142-143 and dataset/coco/get_map.py:57-59;
 *           the /255 of preprocess_input (utils/utils.py:65-67) is folded into the stem weights (tod_stem_conv_nhwc_u8).
 * Pillow's 8-bit resampler is integer arithmetic, so the canvas is reproduced bit for bit.
 *
 * tod_resample_coeffs_bicubic (HOST function, no CUDA): windows and 22-bit fixed-point weights of one axis, Pillow
 *   Resample.c precompute_coeffs + normalize_coeffs_8bpc with the bicubic filter (a = -0.5), full-image box.
 *   With h_bounds == NULL or h_coef == NULL only *ksize is written (size query).
 *     h_bounds i32 [out_size, 2] = (first input index, tap count); h_coef i32 [out_size, ksize] (unused taps 0).
 * tod_letterbox_bicubic_u8: n same-sized uint8 HWC RGB images -> n (dst_h, dst_w, 3) canvases: horizontal pass
 *   (src_w -> new_w, into d_tmp: n * src_h * dst_w * 3 bytes, 4-byte aligned), vertical pass (src_h -> new_h) written at (off_y, off_x) of the
 *   canvas, pad_value everywhere else.  A pass whose size does not change is skipped, like Pillow's.
 *   d_xbounds/d_xcoef/xksize: tod_resample_coeffs_bicubic(src_w, new_w) copied to the device; d_y*: (src_h, new_h).
 */
typedef struct tod_letterbox_desc {
  const uint8_t* d_src;
  uint8_t* d_tmp;
  uint8_t* d_dst;
  int64_t src_image_stride, dst_image_stride; /* bytes between consecutive images */
  int32_t n, src_h, src_w, dst_h, dst_w, new_h, new_w, off_y, off_x, pad_value;
  const int32_t* d_xbounds;
  const int32_t* d_xcoef;
  const int32_t* d_ybounds;
  const int32_t* d_ycoef;
  int32_t xksize, yksize;
  int32_t reserved[4];
} tod_letterbox_desc;
int tod_resample_coeffs_bicubic(int32_t in_size, int32_t out_size, int32_t* h_bounds, int32_t* h_coef, int32_t* ksize);
int tod_letterbox_bicubic_u8(const tod_letterbox_desc* desc, void* stream);

/*
 * CBAM block on NHWC bf16 activations (block-level: not yet wired into the network plan).
 * Replaces: CBAM.forward  model/blocks.py:206-223 (instances: model/backbone.py:26,40, model/head.py:28,30,39,41).
 *   d_x bf16 [batch, h, w, x_pitch] -> d_out bf16 [batch, h, w, out_pitch] (may alias d_x), c channels (c % 8 == 0)
 *   d_fc1 f32 [hidden, c], d_fc2 f32 [c, hidden] (the two bias-free 1x1 convs), d_conv f32 [2, ksize, ksize]
 *   d_work f32, tod_cbam_workspace_floats(batch, h, w, c) elements (pooling partials, channel scale, per-pixel stats)
 */
typedef struct tod_cbam_desc {
  const void* d_x;
  void* d_out;
  const float* d_fc1;
  const float* d_fc2;
  const float* d_conv;
  float* d_work;
  int32_t batch, h, w, c, hidden, ksize, x_pitch, out_pitch;
  int32_t reserved[4];
} tod_cbam_desc;
int64_t tod_cbam_workspace_floats(int32_t batch, int32_t h, int32_t w, int32_t c);
int tod_cbam_nhwc_bf16(const tod_cbam_desc* desc, void* stream);

/*
 * Row softmax f32 -> bf16.
 * Replaces: nn.Softmax(dim=-1) over the key axis in SelfAttention.forward  model/blocks.py:246-247.  The block itself is
 * composed on the host (transparent_object_detection_b200/attention.py) from tod_conv2d_nhwc_bf16 used as a plain
 * GEMM (q / k projections, q . k^T, (gamma Wv) . x^T, P . v^T + gamma b_v + x) and this kernel: an UNFUSED baseline that
 * materialises the N x N score matrix per image; a flash-style tcgen05 kernel is the next step.
 *   d_in f32 [rows, in_pitch] -> d_out bf16 [rows, out_pitch], cols % 4 == 0
 */
int tod_softmax_rows_f32_bf16(const float* d_in, void* d_out, int32_t rows, int32_t cols, int64_t in_pitch,
                              int64_t out_pitch, void* stream);

/*
 * Fused self-attention on tcgen05 / TMEM: the N x N scores never leave the SM.
 * Replaces: torch.bmm(query, key) -> softmax -> torch.bmm(value, attention^T) -> gamma * out + x
 *           SelfAttention.forward  model/blocks.py:243-253 (the three 1x1 projections :239-241 are tod_conv2d_nhwc_bf16 calls).
 *   d_q, d_k  bf16 [batch, n, d16]   query / key projections (+ bias), channels zero-padded to d16 in {16, 32, 64};
 *                                    the QUERY projection (weights and bias) is pre-multiplied by log2(e): the kernel
 *                                    evaluates the softmax with base-2 exponentials
 *   d_vt      bf16 [batch, c, n]     (gamma * Wv) . x^T per image (value projection without bias, transposed: K-major for P . V)
 *   d_bias    f32 [c] = gamma * b_v  (softmax rows sum to one, so the value bias moves out of the sum) or NULL
 *   d_x       bf16 [batch, n, x_pitch] residual; d_out bf16 [batch, n, out_pitch] (may alias d_x); n % 16 == 0, c % 32 == 0, c <= 256
 */
typedef struct tod_attention_desc {
  const void* d_q;
  const void* d_k;
  const void* d_vt;
  const float* d_bias;
  const void* d_x;
  void* d_out;
  int32_t batch, n, c, d16, x_pitch, out_pitch;
  int32_t reserved[4];
} tod_attention_desc;
int tod_attention_fused(const tod_attention_desc* desc, void* stream);
/* [batch, rows, in_pitch] bf16 -> [batch, cols, out_pitch] bf16 transposed per image (the value projection [B, N, C] of the
 * batched 1x1 conv -> V^T [B, C, N] for tod_attention_fused; replaces the .view / .permute of model/blocks.py:241,250). */
int tod_transpose_bf16(const void* d_in, void* d_out, int32_t batch, int32_t rows, int32_t cols, int64_t in_pitch,
                       int64_t out_pitch, void* stream);

/* Debug/verification helper used by tests only: direct (non-tensor-core) evaluation of the same conv
 * descriptor on CUDA cores, fp32 accumulate.  Never called by the product path. */
int tod_conv2d_nhwc_bf16_simt_check(const tod_conv_desc* desc, void* stream);

/* Debug/profiling helper used by tools only: when d_buf is non-NULL the halo conv kernel records per-CTA wait-cycle
 * counters (16 x uint64 per CTA, see conv_halo_tcgen05.cu) into it; NULL switches recording off. */
int tod_debug_set_conv_profile(void* d_buf);

/* Debug/profiling helper used by tools only (tools/timeline.py): while d_buf is non-NULL every instrumented kernel that is
 * ENQUEUED (or captured into a graph) appends one record per CTA {launch id | blockIdx << 32, SM id, %globaltimer at CTA
 * start, at CTA end} to d_buf = u64 cursor, u64 capacity, then 4 x u64 records: which kernels overlap inside the captured
 * graph, and how busy each SM is.  tod_debug_timeline_name(id) names launch `id`; NULL switches it off and resets the ids. */
int tod_debug_set_timeline(void* d_buf);
int tod_debug_timeline_launches(void);
const char* tod_debug_timeline_name(int id);

#ifdef __cplusplus
}
#endif
#endif /* TOD_H_ */

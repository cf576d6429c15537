"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

TEST INFRASTRUCTURE.  Run in the authoring container only:  python -m oracle.make_golden
The fixtures pin oracle/detector_oracle.py (and through it the CUDA path) to the reference's
own arithmetic; inputs/weights are regenerated from seeds by oracle/synth.py, so only outputs
(and a few hand-made adversarial inputs) are stored.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_import, synth  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy()


def ref_nms(ref_bbox, pred: np.ndarray, nc, input_shape, image_shape, letterbox, conf, iou):
    """Run the reference DecodeBox.non_max_suppression; returns (list of arrays|None, mutated prediction)."""
    p = torch.from_numpy(pred.copy())
    db = ref_bbox.DecodeBox(nc, input_shape)
    with torch.no_grad():
        out = db.non_max_suppression(p, nc, input_shape, np.array(image_shape), letterbox, conf_thres=conf, nms_thres=iou)
    return out, p.numpy()


def pack_dets(out):
    """list[None|ndarray(n,6)] -> (concatenated rows, counts with -1 for None)."""
    counts = np.array([-1 if o is None else o.shape[0] for o in out], dtype=np.int64)
    rows = [o for o in out if o is not None]
    rows = np.concatenate(rows, 0).astype(np.float32) if rows else np.zeros((0, 6), np.float32)
    return rows, counts


def adversarial_cases(nc=4):
    """Hand-made (1, A, 4+nc) predictions; boxes given as xyxy then converted to xywh exactly."""
    def mk(boxes_xyxy, scores, classes):
        a = len(boxes_xyxy)
        p = np.zeros((1, a, 4 + nc), np.float32)
        b = np.asarray(boxes_xyxy, np.float32)
        p[0, :, 0] = (b[:, 0] + b[:, 2]) / 2
        p[0, :, 1] = (b[:, 1] + b[:, 3]) / 2
        p[0, :, 2] = b[:, 2] - b[:, 0]
        p[0, :, 3] = b[:, 3] - b[:, 1]
        for i, (s, c) in enumerate(zip(scores, classes)):
            p[0, i, 4 + c] = s
        return p
    cases = {}
    # IoU exactly at the threshold: [0,0,2,2] vs [0,0,2,1] -> 0.5 (kept at thr 0.5)
    cases["iou_eq_thr"] = (mk([[0, 0, 2, 2], [0, 0, 2, 1]], [0.9, 0.8], [1, 1]), 0.5, 0.5)
    # IoU == f32(0.4) with thr 0.4 (double compare suppresses), 2/5
    cases["iou_f32_04"] = (mk([[0, 0, 5, 1], [0, 0, 2, 1]], [0.9, 0.8], [2, 2]), 0.5, 0.4)
    # IoU == f32(0.65) with thr 0.65 (kept), 13/20
    cases["iou_f32_065"] = (mk([[0, 0, 20, 1], [0, 0, 13, 1]], [0.9, 0.8], [0, 0]), 0.5, 0.65)
    # all-equal scores, duplicated boxes, two classes interleaved
    cases["ties_dups"] = (mk([[0, 0, 4, 4]] * 3 + [[1, 1, 5, 5]] * 3 + [[10, 10, 12, 12]] * 2,
                             [0.7] * 8, [3, 1, 3, 1, 3, 1, 3, 1]), 0.5, 0.5)
    # zero-area boxes (NaN IoU survives)
    cases["zero_area"] = (mk([[1, 1, 1, 1], [1, 1, 1, 1], [0, 0, 2, 2], [1, 1, 1, 3]], [0.9, 0.8, 0.7, 0.6], [0, 0, 0, 0]), 0.5, 0.5)
    # class-max tie -> lowest class id; conf exactly at threshold (>=)
    p = mk([[0, 0, 1, 1], [2, 2, 3, 3], [4, 4, 5, 5]], [0.6, 0.5, 0.49999997], [2, 1, 0])
    p[0, 0, 4 + 3] = 0.6  # tie between class 2 and 3
    cases["cls_tie_conf_eq"] = (p, 0.5, 0.5)
    # nothing passes
    cases["empty"] = (mk([[0, 0, 1, 1], [0, 0, 2, 2]], [0.1, 0.2], [0, 1]), 0.5, 0.4)
    # one candidate
    cases["single"] = (mk([[0, 0, 1, 1], [0, 0, 2, 2]], [0.1, 0.9], [0, 1]), 0.5, 0.4)
    return cases


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(8)
    ref_model, ref_bbox = ref_import.import_reference()
    from model import blocks as rb

    # ------------------------------------------------------------------ 1. key table pin
    for name, (C, d, m) in synth.SCALES.items():
        model = ref_import.build_reference_model(80, C, d, m)
        ref_shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        ours = dict(synth.state_dict_shapes(80, C, d, m))
        assert list(ref_shapes.keys()) == list(ours.keys()), f"key order mismatch at scale {name}"
        assert ref_shapes == ours, f"shape mismatch at scale {name}"
        print(f"scale {name}: {len(ours)} keys match, params = "
              f"{sum(int(np.prod(s)) for k, s in ours.items() if 'num_batches' not in k and 'running' not in k)}")

    # ------------------------------------------------------------------ 2. network-level, scale n, 2x3x96x128
    C, d, m = synth.SCALES["n"]
    nc = 80
    sd = synth.make_state_dict(nc, C, d, m, seed=0)
    model = ref_import.build_reference_model(nc, C, d, m, sd)
    x = torch.from_numpy(synth.make_images(2, 96, 128, seed=2))
    with torch.no_grad():
        p3, p4, p5 = model.backbone(x)
        h2, h4, h6 = model.neck((p3, p4, p5))
        model.train()
        # training-mode head on an eval network body: BN stays in eval via functional call below
        model.eval()
        model.head.training = True
        raw = model.head([h2.clone(), h4.clone(), h6.clone()])
        model.head.training = False
        out = model(x)
        # the upstream 5-tuple decode_box == permute + normalise of the head tensor
        xx = torch.cat([r.reshape(2, 64 + nc, -1) for r in raw], 2)
        box, cls = xx.split((64, nc), 1)
        db = ref_bbox.DecodeBox(nc, (96, 128))
        dec = db.decode_box((model.head.dfl(box), cls, None, model.head.anchors, model.head.strides))
        dec2 = out.permute(0, 2, 1).clone()
        dec2[:, :, :4] = dec2[:, :, :4] / torch.tensor([128, 96, 128, 96], dtype=torch.float32)
        assert torch.equal(dec, dec2), "decode_box(5-tuple) != head tensor permute/normalise"
    res = {"p3": _np(p3), "p4": _np(p4), "p5": _np(p5), "h2": _np(h2), "h4": _np(h4), "h6": _np(h6),
           "raw0": _np(raw[0]), "raw1": _np(raw[1]), "raw2": _np(raw[2]), "out": _np(out), "decoded": _np(dec)}
    for conf, iou, tag in ((0.001, 0.65, "coco"), (0.05, 0.5, "cb")):
        o, _ = ref_nms(ref_bbox, _np(dec), nc, (96, 128), (300, 500), True, conf, iou)
        rows, counts = pack_dets(o)
        res[f"nms_{tag}_rows"], res[f"nms_{tag}_counts"] = rows, counts
        print("net_n nms", tag, counts)
    np.savez_compressed(os.path.join(GOLDEN, "net_n_96x128.npz"), **res)

    # ------------------------------------------------------------------ 3. block-level
    g = np.random.Generator(np.random.PCG64(7))
    blk = {}

    def load_conv(mod, prefix, sdict):
        mod.load_state_dict({k[len(prefix) + 1:]: torch.from_numpy(np.asarray(v)) if np.ndim(v) else torch.tensor(v)
                             for k, v in sdict.items() if k.startswith(prefix + ".")})
        return mod.eval()

    def conv_sd(prefix, c1, c2, k, seed):
        t = {}
        synth._conv_keys(t, prefix, c1, c2, k)
        out_ = {}
        for key, shape in t.items():
            r = synth._rng(seed, key)
            if key.endswith("conv.weight"):
                out_[key] = (r.standard_normal(shape) * np.sqrt(2.0 / (shape[1] * k * k))).astype(np.float32)
            elif key.endswith("norm.weight"):
                out_[key] = (1 + 0.02 * r.standard_normal(shape)).astype(np.float32)
            elif key.endswith("running_var"):
                out_[key] = r.uniform(0.5, 1.5, shape).astype(np.float32)
            elif key.endswith("num_batches_tracked"):
                out_[key] = np.zeros((), np.int64)
            else:
                out_[key] = (0.1 * r.standard_normal(shape)).astype(np.float32)
        return out_

    xin = g.standard_normal((2, 16, 10, 12)).astype(np.float32)
    blk["x"] = xin
    xt = torch.from_numpy(xin)
    with torch.no_grad():
        for tag, (c2, k, s) in {"conv1x1": (32, 1, 1), "conv3x3": (16, 3, 1), "conv3x3s2": (32, 3, 2)}.items():
            sdc = conv_sd("c", 16, c2, k, 11)
            mod = load_conv(rb.Conv(16, c2, k, s), "c", sdc)
            blk[tag] = _np(mod(xt))
            fused = rb.fuse_conv(mod.conv, mod.norm)
            blk[tag + "_fw"], blk[tag + "_fb"] = _np(fused.weight), _np(fused.bias)
        # Bottleneck(16,16, shortcut, k=((3,3),(3,3)), e=1.0)
        sdb = {}
        sdb.update(conv_sd("b.cv1", 16, 16, 3, 12)); sdb.update(conv_sd("b.cv2", 16, 16, 3, 12))
        for sc in (True, False):
            mod = load_conv(rb.Bottleneck(16, 16, sc, k=((3, 3), (3, 3)), e=1.0), "b", sdb)
            blk[f"bottleneck_{int(sc)}"] = _np(mod(xt))
        # C2f(16, 32, n=2, shortcut=True)
        sdc2 = {}
        sdc2.update(conv_sd("f.cv1", 16, 32, 1, 13)); sdc2.update(conv_sd("f.cv2", 64, 32, 1, 13))
        for j in range(2):
            sdc2.update(conv_sd(f"f.m.{j}.cv1", 16, 16, 3, 13)); sdc2.update(conv_sd(f"f.m.{j}.cv2", 16, 16, 3, 13))
        mod = load_conv(rb.C2f(16, 32, 2, True), "f", sdc2)
        blk["c2f"] = _np(mod(xt))
        # SPPF(16, 16)
        sds = {}
        sds.update(conv_sd("s.cv1", 16, 8, 1, 14)); sds.update(conv_sd("s.cv2", 32, 16, 1, 14))
        mod = load_conv(rb.SPPF(16, 16, 5), "s", sds)
        blk["sppf"] = _np(mod(xt))
        pools = [xt]
        for _ in range(3):
            pools.append(mod.m(pools[-1]))
        blk["sppf_pools"] = _np(torch.cat(pools, 1))
        # DFL
        logits = (g.standard_normal((2, 64, 50)) * 3).astype(np.float32)
        blk["dfl_in"] = logits
        blk["dfl_out"] = _np(rb.DFL(16)(torch.from_numpy(logits)))
        # make_anchors / dist2bbox
        feats = [torch.zeros(1, 1, 12, 16), torch.zeros(1, 1, 6, 8), torch.zeros(1, 1, 3, 4)]
        a, st = ref_bbox.make_anchors(feats, torch.tensor([8.0, 16.0, 32.0]), 0.5)
        blk["anchors"], blk["anchor_strides"] = _np(a), _np(st)
        dist = np.abs(g.standard_normal((2, 4, 252))).astype(np.float32)
        blk["dist"] = dist
        blk["dist2bbox"] = _np(ref_bbox.dist2bbox(torch.from_numpy(dist), a.transpose(0, 1).unsqueeze(0), xywh=True, dim=1))
        # correct_boxes
        bxy = g.uniform(0.2, 0.8, (5, 2)).astype(np.float32)
        bwh = g.uniform(0.05, 0.3, (5, 2)).astype(np.float32)
        blk["cb_xy"], blk["cb_wh"] = bxy, bwh
        db = ref_bbox.DecodeBox(80, (640, 640))
        blk["cb_letterbox"] = db.correct_boxes(bxy.copy(), bwh.copy(), (640, 640), np.array((375, 500)), True)
        blk["cb_plain"] = db.correct_boxes(bxy.copy(), bwh.copy(), (640, 640), np.array((375, 500)), False)
    np.savez_compressed(os.path.join(GOLDEN, "blocks.npz"), **blk)

    # ------------------------------------------------------------------ 4. NMS cases
    nms = {}
    pred = synth.make_dense_predictions(2, anchors=700, nc=80, objects=24, seed=1234)
    for conf, iou, tag in ((0.001, 0.65, "coco"), (0.05, 0.5, "cb"), (0.5, 0.4, "default")):
        o, mutated = ref_nms(ref_bbox, pred, 80, (640, 640), (480, 640), True, conf, iou)
        nms[f"dense_{tag}_rows"], nms[f"dense_{tag}_counts"] = pack_dets(o)
        print("dense", tag, nms[f"dense_{tag}_counts"])
    nms["dense_mutated_xyxy_sample"] = mutated[:, ::50, :4]
    o, _ = ref_nms(ref_bbox, pred, 80, (640, 640), (480, 640), False, 0.05, 0.5)
    nms["dense_noletterbox_rows"], nms["dense_noletterbox_counts"] = pack_dets(o)
    for name, (p, conf, iou) in adversarial_cases(4).items():
        o, _ = ref_nms(ref_bbox, p, 4, (1, 1), (1, 1), False, conf, iou)
        nms[f"adv_{name}_pred"] = p
        nms[f"adv_{name}_thr"] = np.array([conf, iou], np.float64)
        nms[f"adv_{name}_rows"], nms[f"adv_{name}_counts"] = pack_dets(o)
        print("adv", name, nms[f"adv_{name}_counts"], nms[f"adv_{name}_rows"][:, 4:].tolist())
    # single class, many boxes
    p1 = synth.make_dense_predictions(1, anchors=500, nc=1, objects=10, seed=77)
    o, _ = ref_nms(ref_bbox, p1, 1, (640, 640), (640, 640), False, 0.001, 0.65)
    nms["single_class_rows"], nms["single_class_counts"] = pack_dets(o)
    np.savez_compressed(os.path.join(GOLDEN, "nms_cases.npz"), **nms)

    # ------------------------------------------------------------------ 5. config 1: scale s, 1x3x640x640
    C, d, m = synth.SCALES["s"]
    sd = synth.make_state_dict(80, C, d, m, seed=0)
    model = ref_import.build_reference_model(80, C, d, m, sd)
    x = torch.from_numpy(synth.make_images(1, 640, 640, seed=2))
    with torch.no_grad():
        out = model(x)
    dec = out.permute(0, 2, 1).clone()
    dec[:, :, :4] = dec[:, :, :4] / 640.0
    c1 = {"out_sub": _np(out)[:, :, ::16], "out_sum": np.array(_np(out).astype(np.float64).sum()),
          "score_max": np.array(_np(out)[:, 4:].max())}
    for conf, iou, tag in ((0.001, 0.65, "coco"), (0.05, 0.5, "cb"), (0.5, 0.4, "default")):
        o, _ = ref_nms(ref_bbox, _np(dec), 80, (640, 640), (480, 640), True, conf, iou)
        c1[f"nms_{tag}_rows"], c1[f"nms_{tag}_counts"] = pack_dets(o)
        print("config1 nms", tag, c1[f"nms_{tag}_counts"])
    np.savez_compressed(os.path.join(GOLDEN, "config1_s_640.npz"), **c1)
    print("score max", c1["score_max"])
    for f in sorted(os.listdir(GOLDEN)):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)))


if __name__ == "__main__":
    main()

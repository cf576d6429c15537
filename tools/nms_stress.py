"""BASELINE config 5: dense-detection NMS stress (conf 0.001, IoU 0.65, clustered synthetic boxes): timing of
tod_nms_prepare_dense + tod_nms through DecodeBox.nms_device, keep indices checked against the C oracle for one image.
usage: nms_stress.py [--batch 64] [--anchors 8400] [--objects 120]"""
import argparse, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import detector_oracle as O
from transparent_object_detection_b200 import synth
from transparent_object_detection_b200 import DecodeBox
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64); ap.add_argument("--anchors", type=int, default=8400)
ap.add_argument("--objects", type=int, default=120); ap.add_argument("--conf", type=float, default=0.001); ap.add_argument("--iou", type=float, default=0.65)
a = ap.parse_args()
pred = synth.make_dense_predictions(a.batch, anchors=a.anchors, nc=80, objects=a.objects, seed=1234)
db = DecodeBox(80, (640, 640))
p = torch.from_numpy(pred).cuda()
ts = []
for _ in range(6):
    q = p.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    keep_idx, keep_count, dets = db.nms_device(q, 80, a.conf, a.iou)
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
kc = keep_count.cpu().numpy()
cand = int((pred[:, :, 4:].max(2) >= np.float32(a.conf)).sum())
print(f"batch {a.batch} x {a.anchors} anchors: {cand / a.batch:.0f} candidates/image, {kc.mean():.0f} kept/image; "
      f"{np.median(ts[1:]):.0f} us per batch = {np.median(ts[1:]) / a.batch:.1f} us per image (allocation of the work buffers included)")
want = O.nms_keep_indices(pred[:1].copy(), 80, a.conf, a.iou)[0]
got = keep_idx[0, :kc[0]].cpu().numpy()
print("keep indices of image 0 bit-exact vs the oracle:", bool(np.array_equal(got, np.asarray(want))), f"({kc[0]} kept)")

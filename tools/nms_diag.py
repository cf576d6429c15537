"""Candidate / segment statistics of the bench workload's NMS input and per-kernel NMS timing (CUDA events).
usage: nms_diag.py [--batch 64] [--conf 0.05] [--iou 0.5]"""
import argparse, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transparent_object_detection_b200 import synth
from transparent_object_detection_b200 import BaseModel
ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=64); ap.add_argument("--conf", type=float, default=0.05)
ap.add_argument("--iou", type=float, default=0.5); ap.add_argument("--size", type=int, default=640)
a = ap.parse_args()
C_, d, m = synth.SCALES["s"]
model = BaseModel(80, C_, d, m).eval()
model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.make_state_dict(80, C_, d, m, seed=0).items()})
eng = model.engine(a.batch, a.size, a.size)
x = torch.from_numpy(synth.make_images_u8(a.batch, a.size, a.size, seed=3)).cuda()
eng.run_network(x); eng.run_decode(False, False, True); eng.run_nms(a.conf, a.iou)
torch.cuda.synchronize()
conf = eng.cand_conf.cpu().numpy(); cls = eng.cand_cls.cpu().numpy()
sel = conf >= np.float32(a.conf)
print("candidates/img: mean %.0f max %d" % (sel.sum(1).mean(), sel.sum(1).max()), " kept/img mean %.0f" % eng.keep_count.float().mean().item())
sizes = []
for b in range(a.batch):
    c = cls[b][sel[b]]
    if c.size:
        sizes += list(np.bincount(c, minlength=80)[np.bincount(c, minlength=80) > 0])
sizes = np.array(sizes)
print("segments: %d total, per image %.1f; size mean %.1f median %d p90 %d p99 %d max %d; >256: %d; sum(size^2)=%.3g" % (
    sizes.size, sizes.size / a.batch, sizes.mean(), np.median(sizes), np.percentile(sizes, 90), np.percentile(sizes, 99), sizes.max(),
    (sizes > 256).sum(), float((sizes.astype(np.float64) ** 2).sum())))
for name, fn in (("decode", lambda: eng.run_decode(False, False, True)), ("nms(all 3 kernels)", lambda: eng.run_nms(a.conf, a.iou))):
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    print(f"{name}: {np.median(ts):.1f} us")

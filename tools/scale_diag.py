"""Stage features / head output of one scale against the CPU oracle.  usage: scale_diag.py scale [B H W]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import detector_oracle as O
from transparent_object_detection_b200 import synth
from transparent_object_detection_b200 import BaseModel
scale = sys.argv[1]
B, H, W = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (2, 64, 96)
C_, d, m = synth.SCALES[scale]
sd = synth.make_state_dict(80, C_, d, m, seed=0)
model = BaseModel(80, C_, d, m).eval()
model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
x = torch.from_numpy(synth.make_images(B, H, W, seed=7))
out = model(x.cuda()).cpu()
eng = model.engine(B, H, W, torch.device("cuda", 0))
with torch.no_grad():
    p = O.backbone(sd, x, d)
    h = O.neck(sd, p, d)
    raw = O.head_raw(sd, h)
    want = O.head_decode(raw, 80)
for name, ref in zip(("p3", "p4", "p5", "h2", "h4", "h6"), list(p) + list(h)):
    got = eng.feature_nchw(name).cpu()
    err = (got - ref).abs()
    print(f"{name}: shape {tuple(ref.shape)} max err {float(err.max()):.4f} of absmax {float(ref.abs().max()):.3f}  rms rel {float((err.pow(2).mean().sqrt()) / ref.pow(2).mean().sqrt()):.4f}")
for i, r in enumerate(eng.raw_maps_nchw()):
    e = (r.float().cpu() - raw[i]).abs()
    print(f"raw{i}: max err {float(e.max()):.4f} box part {float(e[:, :64].max()):.4f} cls part {float(e[:, 64:].max()):.4f}  absmax {float(raw[i].abs().max()):.2f}")
be = (out[:, :4] - want[:, :4]).abs(); se = (out[:, 4:] - want[:, 4:]).abs()
print("boxes max err", float(be.max()), "argmax", np.unravel_index(int(be.argmax()), be.shape), "scores max err", float(se.max()), "max score", float(want[:, 4:].max()))

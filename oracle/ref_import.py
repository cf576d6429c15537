"""Import the UNMODIFIED reference from /root/reference as the ground-truth oracle.

TEST INFRASTRUCTURE.  Works only where /root/reference exists (the authoring container);
nothing that runs on the GPU box may import this.  Used by oracle/make_golden.py to produce
tests/golden/*.npz and by the `-m "not gpu"` tests (skipped when the reference is absent).

Accommodations (none touch reference files), SURVEY.md F5/F6/F9 + Appendix C:
  * stub matplotlib / pycocotools into sys.modules (utils/utils_map.py:9-14 would sys.exit);
  * plain topology: backbone.dark{2,3,4}[2] = Identity, neck.h{1,2,4,6} = the reference's own
    C2f(c_in, c_out, depth, False) (channel pairs from model/neck.py:19,25,37,49),
    head.{cls,box}[i][{1,3}] = Identity;
  * head.stride = [8, 16, 32] (model/head.py:17 never sets it).
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "model"))


def _stub(name: str, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference():
    """Returns (model_pkg, bbox_utils module) of the reference."""
    if not available():
        raise RuntimeError("reference not present at " + REFERENCE_ROOT)
    sys.dont_write_bytecode = True  # the mount is read-only
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        mpl = _stub("matplotlib", use=lambda *a, **k: None)
        mpl.pyplot = _stub("matplotlib.pyplot")
    try:
        import pycocotools  # noqa: F401
    except ImportError:
        _stub("pycocotools")
        _stub("pycocotools.coco", COCO=object)
        _stub("pycocotools.cocoeval", COCOeval=object)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import model as ref_model          # the package model/, never model.py
        import utils.bbox_utils as ref_bbox
    return ref_model, ref_bbox


def build_reference_model(nc: int, C: int, d: int, deep_mul: float, state_dict=None, attention: bool = False):
    """Patched-current-source reference BaseModel in eval mode (network-level oracle)."""
    import torch
    import torch.nn as nn
    ref_model, _ = import_reference()
    from model.blocks import C2f
    m = ref_model.BaseModel(nc, C, d, deep_mul)
    C5 = int(C * 16 * deep_mul)
    if not attention:          # attention=True keeps the current source's CBAM / SelfAttention modules
        for name in ("dark2", "dark3", "dark4"):
            getattr(m.backbone, name)[2] = nn.Identity()
    m.neck.h1 = C2f(C5 + 8 * C, 8 * C, d, False)
    m.neck.h2 = C2f(8 * C + 4 * C, 4 * C, d, False)
    m.neck.h4 = C2f(8 * C + 4 * C, 8 * C, d, False)
    m.neck.h6 = C2f(C5 + 8 * C, C5, d, False)
    if not attention:
        for tower in (m.head.cls, m.head.box):
            for seq in tower:
                seq[1] = nn.Identity()
                seq[3] = nn.Identity()
    m.head.stride = torch.tensor([8.0, 16.0, 32.0])
    if state_dict is not None:
        sd = {k: (v if isinstance(v, torch.Tensor) else torch.from_numpy(v.copy()) if v.shape else torch.tensor(v))
              for k, v in state_dict.items()}
        m.load_state_dict(sd, strict=True)
    return m.eval()

"""Fixture for the loss-side decode Loss.bbox_decode (model/loss.py:333-337), written from the reference's
OWN method in the authoring container (/root/reference is absent on the GPU box): tests/golden/loss_bbox_decode.npz.

    python -m oracle.make_golden_loss

Case "net": the box half of the training-mode head maps of the scale-n fixture network, flattened and permuted exactly as
Loss.__call__ does (loss.py:343-347), with make_anchors' points (loss.py:352).  Case "syn": random logits with a large
spread (near one-hot softmax rows) on an anchor count that is not a multiple of anything.  Case "nodfl": reg_max 1.
TEST INFRASTRUCTURE."""
from __future__ import annotations

import os
import types

import numpy as np
import torch

from . import ref_import, synth

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    torch.manual_seed(0)
    ref_import.import_reference()
    from model.loss import Loss            # the reference's class; bbox_decode only reads use_dfl and proj
    from utils.bbox_utils import make_anchors
    C, d, m = synth.SCALES["n"]
    nc = 80
    sd = synth.make_state_dict(nc, C, d, m, seed=0)
    model = ref_import.build_reference_model(nc, C, d, m, sd)
    x = torch.from_numpy(synth.make_images(2, 96, 128, seed=2))
    res = {}
    with torch.no_grad():
        model.head.training = True
        feats = model.head(list(model.neck(model.backbone(x))))         # raw (B, 64 + nc, h, w) maps, model/head.py:50-51
        model.head.training = False
        no = 64 + nc
        pred_distri, _ = torch.cat([xi.view(feats[0].shape[0], no, -1) for xi in feats], 2).split((64, nc), 1)   # loss.py:343
        pred_distri = pred_distri.permute(0, 2, 1).contiguous()                                                 # loss.py:347
        anchor_points, _ = make_anchors(feats, [8, 16, 32], 0.5)                                               # loss.py:352
        me = types.SimpleNamespace(use_dfl=True, proj=torch.arange(16, dtype=torch.float))
        out = Loss.bbox_decode(me, anchor_points, pred_distri)
        res.update(net_pred_dist=pred_distri.numpy(), net_anchor_points=anchor_points.numpy(), net_boxes=out.numpy())
        g = torch.Generator().manual_seed(21)
        A = 333
        pd2 = torch.randn((3, A, 64), generator=g) * 8
        ap2 = torch.rand((A, 2), generator=g) * 80
        res.update(syn_pred_dist=pd2.numpy(), syn_anchor_points=ap2.numpy(), syn_boxes=Loss.bbox_decode(me, ap2, pd2).numpy())
        me1 = types.SimpleNamespace(use_dfl=False, proj=torch.arange(1, dtype=torch.float))
        pd3 = torch.rand((2, 50, 4), generator=g) * 10
        ap3 = torch.rand((50, 2), generator=g) * 20
        res.update(nodfl_pred_dist=pd3.numpy(), nodfl_anchor_points=ap3.numpy(), nodfl_boxes=Loss.bbox_decode(me1, ap3, pd3).numpy())
    np.savez_compressed(os.path.join(GOLDEN, "loss_bbox_decode.npz"), **res)
    print("wrote loss_bbox_decode.npz:", {k: v.shape for k, v in res.items()})


if __name__ == "__main__":
    main()

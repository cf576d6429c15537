"""Fixture for DecodeBox.decode_box on the upstream 5-tuple (utils/bbox_utils.py:66-82), written from the reference's OWN
code in the authoring container (/root/reference is absent on the GPU box): tests/golden/decode_tuple.npz.

    python -m oracle.make_golden_tuple

Inputs are what the reference's callers would hand over (utils/callbacks.py:150-151): the DFL distances and class logits
of the reference head on the scale-n fixture network, its anchors / strides (model/head.py:53), plus a second synthetic
case with large logits and distances (saturating sigmoids, boxes beyond the image).  TEST INFRASTRUCTURE."""
from __future__ import annotations

import os

import numpy as np
import torch

from . import ref_import, synth

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    torch.manual_seed(0)
    ref_model, ref_bbox = ref_import.import_reference()
    C, d, m = synth.SCALES["n"]
    nc = 80
    sd = synth.make_state_dict(nc, C, d, m, seed=0)
    model = ref_import.build_reference_model(nc, C, d, m, sd)
    x = torch.from_numpy(synth.make_images(2, 96, 128, seed=2))
    res = {}
    with torch.no_grad():
        feats = list(model.neck(model.backbone(x)))
        model.head.training = True
        raw = model.head([f.clone() for f in feats])
        model.head.training = False
        out = model(x)                                   # sets head.anchors / head.strides (model/head.py:53)
        xx = torch.cat([r.reshape(2, 64 + nc, -1) for r in raw], 2)
        box, cls = xx.split((64, nc), 1)
        dbox = model.head.dfl(box)
        db = ref_bbox.DecodeBox(nc, (96, 128))
        dec = db.decode_box((dbox, cls, None, model.head.anchors, model.head.strides))
        assert torch.equal(dec[:, :, 4:], out.permute(0, 2, 1)[:, :, 4:])
        res.update(net_dbox=dbox.numpy(), net_cls=cls.numpy(), net_anchors=model.head.anchors.contiguous().numpy(),
                   net_strides=model.head.strides.contiguous().numpy(), net_decoded=dec.numpy(), net_input_shape=np.array([96, 128]))
        # synthetic: A not a multiple of 32, one class, extreme values
        g = torch.Generator().manual_seed(11)
        A, nc2 = 77, 3
        dbox2 = torch.rand((3, 4, A), generator=g) * 40 - 5
        cls2 = torch.randn((3, nc2, A), generator=g) * 30
        anchors2 = torch.rand((2, A), generator=g) * 50
        strides2 = torch.tensor([8.0, 16.0, 32.0])[torch.randint(0, 3, (A,), generator=g)].view(1, A)
        db2 = ref_bbox.DecodeBox(nc2, (160, 224))
        dec2 = db2.decode_box((dbox2, cls2, None, anchors2, strides2))
        res.update(syn_dbox=dbox2.numpy(), syn_cls=cls2.numpy(), syn_anchors=anchors2.numpy(), syn_strides=strides2.numpy(),
                   syn_decoded=dec2.numpy(), syn_input_shape=np.array([160, 224]))
    np.savez_compressed(os.path.join(GOLDEN, "decode_tuple.npz"), **res)
    print("wrote decode_tuple.npz:", {k: v.shape for k, v in res.items()})


if __name__ == "__main__":
    main()

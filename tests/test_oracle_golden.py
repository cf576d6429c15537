"""CPU: pin oracle/detector_oracle.py to fixtures produced by the reference itself
(oracle/make_golden.py).  Float stages: tight tolerance (the GPU box's CPU may pick other
oneDNN kernels than the authoring container's); NMS rows/indices on stored inputs: bit-exact."""
import numpy as np
import pytest
import torch

from oracle import detector_oracle as O
from oracle import synth


def unpack(rows, counts):
    out, p = [], 0
    for c in counts:
        if c < 0:
            out.append(None)
        else:
            out.append(rows[p:p + c]); p += c
    return out


def assert_dets_equal(got, want):
    assert len(got) == len(want)
    for g, w in zip(got, want):
        if w is None:
            assert g is None
        else:
            assert g is not None and g.dtype == np.float32 and g.shape == w.shape
            assert np.array_equal(g, w)


def test_state_dict_table_counts():
    # 3 157 200 params at scale n, 11.167 M at scale s
    for scale, want in (("n", 3157200), ("s", 11166560)):
        C, d, m = synth.SCALES[scale]
        t = synth.state_dict_shapes(80, C, d, m)
        n = sum(int(np.prod(s)) for k, s in t.items() if "running" not in k and "num_batches" not in k)
        assert n == want
        assert len(t) == 355


def test_network_stages_match_reference(golden):
    g = golden("net_n_96x128.npz")
    C, d, m = synth.SCALES["n"]
    sd = synth.make_state_dict(80, C, d, m, seed=0)
    x = torch.from_numpy(synth.make_images(2, 96, 128, seed=2))
    with torch.no_grad():
        feats = O.backbone(sd, x, d)
        nk = O.neck(sd, feats, d)
        raw = O.head_raw(sd, nk)
        out = O.head_decode(raw, 80)
        dec = O.decode_box(out, (96, 128))
    for name, t in zip(("p3", "p4", "p5", "h2", "h4", "h6", "raw0", "raw1", "raw2"), (*feats, *nk, *raw)):
        np.testing.assert_allclose(t.numpy(), g[name], rtol=1e-4, atol=1e-5, err_msg=name)
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(dec.numpy(), g["decoded"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("tag,conf,iou", [("coco", 0.001, 0.65), ("cb", 0.05, 0.5)])
def test_network_nms_bit_exact_on_reference_decoded(golden, tag, conf, iou):
    g = golden("net_n_96x128.npz")
    got = O.non_max_suppression(g["decoded"].copy(), 80, (96, 128), (300, 500), True, conf, iou)
    assert_dets_equal(got, unpack(g[f"nms_{tag}_rows"], g[f"nms_{tag}_counts"]))


def test_blocks_match_reference(golden):
    g = golden("blocks.npz")
    # DFL, anchors, dist2bbox, correct_boxes are closed-form restatements
    np.testing.assert_allclose(O.dfl(torch.from_numpy(g["dfl_in"])).numpy(), g["dfl_out"], rtol=1e-5, atol=1e-5)
    a, st = O.make_anchors([(12, 16), (6, 8), (3, 4)], (8.0, 16.0, 32.0))
    assert np.array_equal(a.numpy(), g["anchors"]) and np.array_equal(st.numpy(), g["anchor_strides"])
    x = torch.from_numpy(g["x"])
    np.testing.assert_array_equal(O.sppf_pools(x).numpy(), g["sppf_pools"])
    for lb, key in ((True, "cb_letterbox"), (False, "cb_plain")):
        got = O.correct_boxes(g["cb_xy"].copy(), g["cb_wh"].copy(), (640, 640), (375, 500), lb)
        assert got.dtype == g[key].dtype
        np.testing.assert_array_equal(got, g[key])


def test_fold_bn_matches_reference_fuse_conv(golden):
    g = golden("blocks.npz")
    # regenerate the same block weights as make_golden.conv_sd (seed 11)
    for tag, (c2, k) in {"conv1x1": (32, 1), "conv3x3": (16, 3), "conv3x3s2": (32, 3)}.items():
        t = {}
        synth._conv_keys(t, "c", 16, c2, k)
        sd = {}
        for key, shape in t.items():
            r = synth._rng(11, key)
            if key.endswith("conv.weight"):
                sd[key] = (r.standard_normal(shape) * np.sqrt(2.0 / (shape[1] * k * k))).astype(np.float32)
            elif key.endswith("norm.weight"):
                sd[key] = (1 + 0.02 * r.standard_normal(shape)).astype(np.float32)
            elif key.endswith("running_var"):
                sd[key] = r.uniform(0.5, 1.5, shape).astype(np.float32)
            elif key.endswith("num_batches_tracked"):
                sd[key] = np.zeros((), np.int64)
            else:
                sd[key] = (0.1 * r.standard_normal(shape)).astype(np.float32)
        wf, bf = O.fold_bn(sd, "c")
        np.testing.assert_allclose(wf.numpy(), g[tag + "_fw"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(bf.numpy(), g[tag + "_fb"], rtol=1e-6, atol=1e-7)
        stride = 2 if tag.endswith("s2") else 1
        y = O.conv_bn_silu(sd, "c", torch.from_numpy(g["x"]), stride)
        np.testing.assert_allclose(y.numpy(), g[tag], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("tag,conf,iou", [("coco", 0.001, 0.65), ("cb", 0.05, 0.5), ("default", 0.5, 0.4)])
def test_dense_nms_bit_exact(golden, tag, conf, iou):
    g = golden("nms_cases.npz")
    pred = synth.make_dense_predictions(2, anchors=700, nc=80, objects=24, seed=1234)
    got = O.non_max_suppression(pred, 80, (640, 640), (480, 640), True, conf, iou)
    assert_dets_equal(got, unpack(g[f"dense_{tag}_rows"], g[f"dense_{tag}_counts"]))
    # in-place xywh -> xyxy side effect (utils/bbox_utils.py:144-149)
    assert np.array_equal(pred[:, ::50, :4], g["dense_mutated_xyxy_sample"])


def test_dense_nms_no_letterbox(golden):
    g = golden("nms_cases.npz")
    pred = synth.make_dense_predictions(2, anchors=700, nc=80, objects=24, seed=1234)
    got = O.non_max_suppression(pred, 80, (640, 640), (480, 640), False, 0.05, 0.5)
    assert_dets_equal(got, unpack(g["dense_noletterbox_rows"], g["dense_noletterbox_counts"]))


def test_adversarial_nms_bit_exact(golden):
    g = golden("nms_cases.npz")
    names = sorted({k[4:-5] for k in g.files if k.startswith("adv_") and k.endswith("_pred")})
    assert len(names) == 8
    for name in names:
        conf, iou = g[f"adv_{name}_thr"]
        got = O.non_max_suppression(g[f"adv_{name}_pred"].copy(), 4, (1, 1), (1, 1), False, float(conf), float(iou))
        assert_dets_equal(got, unpack(g[f"adv_{name}_rows"], g[f"adv_{name}_counts"]))


def test_single_class_nms(golden):
    g = golden("nms_cases.npz")
    p1 = synth.make_dense_predictions(1, anchors=500, nc=1, objects=10, seed=77)
    got = O.non_max_suppression(p1, 1, (640, 640), (640, 640), False, 0.001, 0.65)
    assert_dets_equal(got, unpack(g["single_class_rows"], g["single_class_counts"]))


def test_config1_scale_s_640(golden):
    g = golden("config1_s_640.npz")
    C, d, m = synth.SCALES["s"]
    sd = synth.make_state_dict(80, C, d, m, seed=0)
    x = torch.from_numpy(synth.make_images(1, 640, 640, seed=2))
    with torch.no_grad():
        out = O.forward(sd, x, 80, d)
    assert out.shape == (1, 84, 8400)
    np.testing.assert_allclose(out.numpy()[:, :, ::16], g["out_sub"], rtol=1e-4, atol=2e-4)
    np.testing.assert_allclose(out.numpy().astype(np.float64).sum(), float(g["out_sum"]), rtol=1e-6)


# ---------------------------------------------------------------- plain-C restatement (oracle/nms_oracle.c)
@pytest.mark.parametrize("tag,conf,iou", [("coco", 0.001, 0.65), ("cb", 0.05, 0.5), ("default", 0.5, 0.4)])
def test_c_oracle_dense_nms_bit_exact(golden, tag, conf, iou):
    from oracle import nms_c
    g = golden("nms_cases.npz")
    pred = synth.make_dense_predictions(2, anchors=700, nc=80, objects=24, seed=1234)
    got = O.non_max_suppression(pred, 80, (640, 640), (480, 640), True, conf, iou, keep_fn=nms_c.nms_keep_indices)
    assert_dets_equal(got, unpack(g[f"dense_{tag}_rows"], g[f"dense_{tag}_counts"]))


def test_c_oracle_adversarial_nms_bit_exact(golden):
    from oracle import nms_c
    g = golden("nms_cases.npz")
    names = sorted({k[4:-5] for k in g.files if k.startswith("adv_") and k.endswith("_pred")})
    for name in names:
        conf, iou = g[f"adv_{name}_thr"]
        got = O.non_max_suppression(g[f"adv_{name}_pred"].copy(), 4, (1, 1), (1, 1), False, float(conf), float(iou),
                                    keep_fn=nms_c.nms_keep_indices)
        assert_dets_equal(got, unpack(g[f"adv_{name}_rows"], g[f"adv_{name}_counts"]))


def test_c_oracle_matches_numpy_oracle_at_full_size():
    from oracle import nms_c
    pred = synth.make_dense_predictions(2, anchors=8400, nc=80, objects=120, seed=1234)
    a = nms_c.nms_keep_indices(pred, 80, 0.001, 0.65)
    b = O.nms_keep_indices(pred, 80, 0.001, 0.65)
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and len(a[0]) > 100


@pytest.mark.parametrize("case", ["net", "syn"])
def test_decode_box_tuple_oracle_bit_exact_vs_reference(golden, case):
    """oracle.decode_box_tuple == the reference's DecodeBox.decode_box on the upstream 5-tuple (utils/bbox_utils.py:66-82),
    fixture written by oracle/make_golden_tuple.py from the reference's own code."""
    g = golden("decode_tuple.npz")
    t = lambda k: torch.from_numpy(g[f"{case}_{k}"])
    got = O.decode_box_tuple(t("dbox"), t("cls"), t("anchors"), t("strides"), tuple(int(v) for v in g[f"{case}_input_shape"]))
    want = g[f"{case}_decoded"]
    assert np.array_equal(got.numpy()[:, :, :4], want[:, :, :4])          # add / sub / mul / div only: bit-exact anywhere
    assert np.abs(got.numpy()[:, :, 4:] - want[:, :, 4:]).max() <= 1e-6   # sigmoid: the host's vector exp may differ by ulps


def test_oracle_loss_bbox_decode_matches_reference_fixture(golden):
    """the oracle's restatement of Loss.bbox_decode (model/loss.py:333-337) against the fixture written from
    the reference's own method (oracle/make_golden_loss.py) -- same torch ops, so bit for bit."""
    g = golden("loss_bbox_decode.npz")
    for case, reg_max in (("net", 16), ("syn", 16), ("nodfl", 1)):
        got = O.loss_bbox_decode(torch.from_numpy(g[case + "_anchor_points"]), torch.from_numpy(g[case + "_pred_dist"]), reg_max)
        assert np.array_equal(got.numpy(), g[case + "_boxes"]), case

"""GPU parity of the non-GEMM kernels and of the whole path (through the C ABI / drop-in classes) against the
CPU oracle (oracle/detector_oracle.py) and the reference-generated fixtures in tests/golden/.

Stated tolerances (SURVEY.md section 8c, bf16 activations with fp32 accumulation; scales n and s):
  stage features : max-abs error <= 2 % of the reference abs-max, RMS-relative <= 1 %
  decoded boxes  : <= 1 px;  scores <= 5e-3
  decode kernel on fp32 raw maps: boxes <= 2e-3 px, scores <= 2e-6 (fp32 exp differences only)
  NMS            : keep indices / rows bit-exact when fed the oracle's (reference's) decoded tensor
"""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _engine_parts():
    from transparent_object_detection_b200 import _lib
    return _lib, _lib.lib()


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
    for i, (g, w) in enumerate(zip(got, want)):
        if w is None:
            assert g is None, f"image {i}: expected None"
        else:
            assert g is not None and g.shape == w.shape, f"image {i}: {None if g is None else g.shape} vs {w.shape}"
            assert g.dtype == np.float32
            assert np.array_equal(g, w), f"image {i}: rows differ (max abs {np.abs(g - w).max()})"


# ------------------------------------------------------------------------------------------ stem
@pytest.mark.parametrize("B,H,W,cout", [(2, 64, 96, 32), (1, 32, 32, 16), (1, 640, 640, 32), (1, 64, 64, 96)])
def test_stem_conv(B, H, W, cout):
    _lib, L = _engine_parts()
    g = torch.Generator().manual_seed(1)
    x = torch.rand((B, 3, H, W), generator=g).cuda()
    w = torch.randn((cout, 3, 3, 3), generator=g) * (2.0 / 27) ** 0.5          # host: weights are a kernel parameter
    b = torch.randn((cout,), generator=g) * 0.3
    out = torch.zeros((B, H // 2, W // 2, cout), dtype=torch.bfloat16).cuda()
    wh = w.reshape(cout, 27).contiguous()
    _lib.check(L.tod_stem_conv_nchw_f32(x.data_ptr(), wh.data_ptr(), b.data_ptr(),
                                        out.data_ptr(), B, H, W, cout, cout, torch.cuda.current_stream().cuda_stream), "stem")
    want = F.silu(F.conv2d(x, w.cuda(), b.cuda(), stride=2, padding=1)).permute(0, 2, 3, 1)
    err = (out.float() - want).abs()
    assert float((err - 1e-2 * want.abs()).max()) <= 1e-2, float(err.max())


# ------------------------------------------------------------------------------------------ SPPF pooling
@pytest.mark.parametrize("B,H,W,c", [(2, 20, 20, 256), (1, 40, 40, 256), (1, 3, 4, 16), (1, 7, 5, 24), (1, 80, 80, 16)])
def test_sppf_pool_exact(B, H, W, c):
    from oracle import detector_oracle as O
    _lib, L = _engine_parts()
    g = torch.Generator().manual_seed(2)
    buf = torch.zeros((B, H, W, 4 * c), dtype=torch.bfloat16)
    buf[..., :c] = torch.randn((B, H, W, c), generator=g).to(torch.bfloat16)
    want = O.sppf_pools(buf[..., :c].float().permute(0, 3, 1, 2)).permute(0, 2, 3, 1)   # max is exact in bf16
    d = buf.cuda()
    _lib.check(L.tod_sppf_pool_nhwc_bf16(d.data_ptr(), B, H, W, c, 4 * c, torch.cuda.current_stream().cuda_stream), "pool")
    assert torch.equal(d.float().cpu(), want)


# ------------------------------------------------------------------------------------------ decode
def _run_decode(raw_nchw, nc, in_h, in_w, full=True):
    """raw_nchw: list of 3 cpu f32 (B, 64+nc, h, w) -> dict of device outputs.  full=False: NMS candidates only (the
    kernel's logit-max fast path)."""
    _lib, L = _engine_parts()
    B = raw_nchw[0].shape[0]
    pitch = 64 + (nc + 15) // 16 * 16
    d = _lib.DecodeDesc()
    keep, A = [], 0
    for i, r in enumerate(raw_nchw):
        h, w = r.shape[2:]
        t = torch.zeros((B, h, w, pitch), dtype=torch.float32)
        t[..., :64 + nc] = r.permute(0, 2, 3, 1)
        t = t.cuda()
        keep.append(t)
        d.d_raw[i], d.h[i], d.w[i], d.stride[i] = t.data_ptr(), h, w, float(in_h // h)
        A += h * w
    out = dict(head=torch.zeros((B, 4 + nc, A)).cuda(), dec=torch.zeros((B, A, 4 + nc)).cuda(),
               box=torch.zeros((B, A, 4)).cuda(), conf=torch.zeros((B, A)).cuda(),
               cls=torch.zeros((B, A), dtype=torch.int32).cuda())
    d.raw_pitch, d.batch, d.nc, d.in_h, d.in_w = pitch, B, nc, in_h, in_w
    if full:
        d.d_head_out, d.d_decoded = out["head"].data_ptr(), out["dec"].data_ptr()
    d.d_cand_box, d.d_cand_conf, d.d_cand_cls = out["box"].data_ptr(), out["conf"].data_ptr(), out["cls"].data_ptr()
    _lib.check(L.tod_head_decode(C.byref(d), torch.cuda.current_stream().cuda_stream), "decode")
    torch.cuda.synchronize()
    return out


def test_head_decode_matches_reference_fixture(golden):
    from oracle import detector_oracle as O
    g = golden("net_n_96x128.npz")
    raw = [torch.from_numpy(g[f"raw{i}"]) for i in range(3)]
    out = _run_decode(raw, 80, 96, 128)
    want_head, want_dec = g["out"], g["decoded"]
    got_head = out["head"].cpu().numpy()
    assert np.abs(got_head[:, :4] - want_head[:, :4]).max() <= 2e-3
    assert np.abs(got_head[:, 4:] - want_head[:, 4:]).max() <= 2e-6
    got_dec = out["dec"].cpu().numpy()
    assert np.abs(got_dec[:, :, :4] - want_dec[:, :, :4]).max() <= 2e-5
    assert np.abs(got_dec[:, :, 4:] - want_dec[:, :, 4:]).max() <= 2e-6
    # candidates are self-consistent with the kernel's own decoded tensor, bit for bit
    dec = out["dec"].cpu()
    half = dec[:, :, 2:4] / 2
    corners = torch.cat((dec[:, :, 0:2] - half, dec[:, :, 0:2] + half), 2)
    assert torch.equal(out["box"].cpu(), corners)
    conf, cls = dec[:, :, 4:].max(2)
    assert torch.equal(out["conf"].cpu(), conf)
    assert torch.equal(out["cls"].cpu().long(), cls)


def test_decode_box_from_head_bit_exact(golden):
    from transparent_object_detection_b200 import DecodeBox
    g = golden("net_n_96x128.npz")
    db = DecodeBox(80, (96, 128))
    got = db.decode_box(torch.from_numpy(g["out"]).cuda())
    assert np.array_equal(got.cpu().numpy(), g["decoded"])


@pytest.mark.parametrize("case", ["net", "syn"])
@pytest.mark.parametrize("on_cpu", [False, True])
def test_decode_box_tuple_form_vs_reference_fixture(golden, case, on_cpu):
    """DecodeBox.decode_box on the upstream 5-tuple (dbox, cls, origin_cls, anchors, strides), utils/bbox_utils.py:66-82:
    box columns bit-exact with the reference, scores within 2e-6 (fp32 exp), inputs given on either device."""
    from transparent_object_detection_b200 import DecodeBox
    g = golden("decode_tuple.npz")
    t = lambda k: torch.from_numpy(g[f"{case}_{k}"]) if on_cpu else torch.from_numpy(g[f"{case}_{k}"]).cuda()
    want = g[f"{case}_decoded"]
    nc = want.shape[2] - 4
    db = DecodeBox(nc, tuple(int(v) for v in g[f"{case}_input_shape"]))
    anchors = t("anchors").t().contiguous().t()            # the reference hands over a transposed VIEW (model/head.py:53)
    got = db.decode_box((t("dbox"), t("cls"), None, anchors, t("strides")))
    assert got.is_cuda != on_cpu and tuple(got.shape) == want.shape
    got = got.cpu().numpy()
    assert np.array_equal(got[:, :, :4], want[:, :, :4])
    assert np.abs(got[:, :, 4:] - want[:, :, 4:]).max() <= 2e-6
    with pytest.raises(ValueError):
        db.decode_box((t("dbox"), t("cls"), None))


def test_decode_box_tuple_equals_head_tensor_form(golden):
    """decode_box(5-tuple) == decode_box(head tensor) on the same logits -- through the two C-ABI entries."""
    from transparent_object_detection_b200 import DecodeBox
    g, gt = golden("net_n_96x128.npz"), golden("decode_tuple.npz")
    db = DecodeBox(80, (96, 128))
    a = db.decode_box(tuple(torch.from_numpy(gt[f"net_{k}"]).cuda() if k else None for k in ("dbox", "cls", None, "anchors", "strides")))
    b = db.decode_box(torch.from_numpy(g["out"]).cuda())
    assert torch.equal(a[:, :, :4], b[:, :, :4])
    assert float((a[:, :, 4:] - b[:, :, 4:]).abs().max()) <= 2e-6


def test_decode_nc1_ragged_tiles():
    """nc = 1 (the reference's own coco_classes.txt) and level sizes that are not multiples of the CTA tile."""
    from oracle import detector_oracle as O
    g = torch.Generator().manual_seed(3)
    raw = [torch.randn((2, 65, h, w), generator=g) * 2 for h, w in ((12, 20), (6, 10), (3, 5))]
    out = _run_decode(raw, 1, 96, 160)
    want = O.head_decode(raw, 1).numpy()
    got = out["head"].cpu().numpy()
    assert np.abs(got[:, :4] - want[:, :4]).max() <= 2e-3 and np.abs(got[:, 4:] - want[:, 4:]).max() <= 2e-6
    assert int(out["cls"].abs().max()) == 0


@pytest.mark.parametrize("nc", [80, 1, 3, 21, 200])
def test_decode_candidates_only_path_equals_full_path_on_adversarial_logits(nc):
    """The candidates-only kernel finds the class maximum on the logits and evaluates the sigmoid only inside a tie
    window; it must give bit-identical (box, conf, cls) to the full path, whose conf/cls are checked against torch.max
    over the kernel's own score tensor (lowest class wins ties): exact ties, near ties below float32 score resolution,
    saturated logits (score 1.0 for many classes), hugely negative logits (score 0 for all) and +-inf."""
    g = torch.Generator().manual_seed(11)
    shapes = ((12, 20), (6, 10), (3, 5))
    raw = [torch.randn((2, 64 + nc, h, w), generator=g) * 3 for h, w in shapes]
    r0 = raw[0]
    cls0 = r0[:, 64:]
    if nc >= 3:
        cls0[0, :, 0, 0] = -4.0; cls0[0, 2, 0, 0] = 1.5; cls0[0, 1, 0, 0] = 1.5          # exact tie: class 1 wins
        cls0[0, :, 0, 1] = 20.0 + torch.arange(nc, dtype=torch.float32)                   # all saturate to 1.0: class 0
        cls0[0, :, 0, 2] = -200.0                                                         # all scores 0: class 0
        cls0[0, :, 0, 3] = 12.0; cls0[0, nc - 1, 0, 3] = 12.0000095                       # near tie, scores may collide
        cls0[0, :, 0, 4] = 9.0; cls0[0, nc - 1, 0, 4] = 9.9                               # inside the wide window
        cls0[0, :, 0, 5] = -3.0; cls0[0, nc - 1, 0, 5] = -3.0 + 2e-7                      # adjacent floats
        cls0[0, :, 0, 6] = float("-inf"); cls0[0, 1, 0, 6] = -90.0
        cls0[0, :, 0, 7] = 3.0; cls0[0, nc // 2, 0, 7] = float("inf")
        cls0[1, :, 1, :] = torch.randint(-2, 3, (nc, 20), generator=g).float()            # many exact ties
        # sweep around the tie window (1e-4 below logit 2, 0.01 up to 8, 2.0 above): best logit m, runner-up m - gap, the
        # runner-up alternately in a lower / higher class; gaps from one float32 step to just outside the window
        ms = [-79.0, -40.0, -12.0, -5.0, -1.0, 0.0, 1.0, 1.99, 2.01, 4.0, 7.9, 8.1, 11.0, 14.9]
        gaps = [0.0, 1.0, 3.0, 5e-5, 0.99e-4, 1.01e-4, 2e-4, 9e-3, 1.1e-2, 1.9, 2.1]      # 1.0 / 3.0 = that many ulps of m
        k = 0
        for m_ in ms:
            for gp in gaps:
                yy, xx = 2 + k // 20, k % 20
                if yy >= 12:
                    break
                mt = torch.tensor(m_, dtype=torch.float32)
                if gp in (1.0, 3.0):
                    lo = mt.clone()
                    for _ in range(int(gp)):
                        lo = torch.nextafter(lo, torch.tensor(-float("inf")))
                else:
                    lo = mt - gp
                cls0[1, :, yy, xx] = m_ - 3.0
                a, b = (1, nc - 1) if k % 2 == 0 else (nc - 1, 1)
                cls0[1, a, yy, xx] = mt
                cls0[1, b, yy, xx] = lo
                k += 1
    full = _run_decode(raw, nc, 96, 160, full=True)
    fast = _run_decode(raw, nc, 96, 160, full=False)
    for k in ("box", "conf", "cls"):
        assert torch.equal(full[k], fast[k]), k
    conf, cls = full["dec"][:, :, 4:].max(2)
    assert torch.equal(full["conf"], conf)
    assert torch.equal(full["cls"].long(), cls)


# ------------------------------------------------------------------------------------------ NMS
def _nms_api(pred_np, nc, input_shape, image_shape, letterbox, conf, iou):
    from transparent_object_detection_b200 import DecodeBox
    p = torch.from_numpy(pred_np.copy()).cuda()
    out = DecodeBox(nc, input_shape).non_max_suppression(p, nc, input_shape, np.array(image_shape), letterbox,
                                                          conf_thres=conf, nms_thres=iou)
    return out, p.cpu().numpy()


@pytest.mark.parametrize("tag,conf,iou", [("coco", 0.001, 0.65), ("cb", 0.05, 0.5), ("default", 0.5, 0.4)])
def test_nms_dense_bit_exact_vs_reference_fixture(golden, tag, conf, iou):
    from oracle import synth
    g = golden("nms_cases.npz")
    pred = synth.make_dense_predictions(2, anchors=700, nc=80, objects=24, seed=1234)
    got, mutated = _nms_api(pred, 80, (640, 640), (480, 640), True, conf, iou)
    assert_dets_equal(got, unpack(g[f"dense_{tag}_rows"], g[f"dense_{tag}_counts"]))
    assert np.array_equal(mutated[:, ::50, :4], g["dense_mutated_xyxy_sample"])     # in-place side effect


def test_nms_adversarial_bit_exact(golden):
    g = golden("nms_cases.npz")
    names = sorted({k[4:-5] for k in g.files if k.startswith("adv_") and k.endswith("_pred")})
    for name in names:
        conf, iou = g[f"adv_{name}_thr"]
        got, _ = _nms_api(g[f"adv_{name}_pred"], 4, (1, 1), (1, 1), False, float(conf), float(iou))
        try:
            assert_dets_equal(got, unpack(g[f"adv_{name}_rows"], g[f"adv_{name}_counts"]))
        except AssertionError as e:
            raise AssertionError(f"case {name}: {e}")


def test_nms_single_class(golden):
    from oracle import synth
    g = golden("nms_cases.npz")
    p1 = synth.make_dense_predictions(1, anchors=500, nc=1, objects=10, seed=77)
    got, _ = _nms_api(p1, 1, (640, 640), (640, 640), False, 0.001, 0.65)
    assert_dets_equal(got, unpack(g["single_class_rows"], g["single_class_counts"]))


@pytest.mark.parametrize("B,A,nc,conf,iou", [(4, 8400, 80, 0.001, 0.65), (2, 8400, 80, 0.25, 0.45), (1, 33600, 80, 0.001, 0.65),
                                             (2, 3000, 1, 0.001, 0.65), (3, 1000, 7, 0.0, 0.3)])
def test_nms_keep_indices_bit_exact_vs_oracle(B, A, nc, conf, iou):
    """Config 5 (dense NMS stress) at full size: keep indices in reference order, bit-exact."""
    from oracle import detector_oracle as O, synth
    from transparent_object_detection_b200 import DecodeBox
    pred = synth.make_dense_predictions(B, anchors=A, nc=nc, objects=120, seed=1234 + A)
    want = O.nms_keep_indices(pred, nc, conf, iou)
    p = torch.from_numpy(pred.copy()).cuda()
    keep_idx, keep_count, dets = DecodeBox(nc, (640, 640)).nms_device(p, nc, conf, iou)
    keep_idx, keep_count = keep_idx.cpu().numpy(), keep_count.cpu().numpy()
    for b in range(B):
        assert keep_count[b] == len(want[b]), (b, keep_count[b], len(want[b]))
        assert np.array_equal(keep_idx[b, :keep_count[b]], want[b]), b
    # size-independent properties: kept rows sorted (class asc, score desc), and idempotence
    d = dets.cpu().numpy()
    for b in range(B):
        r = d[b, :keep_count[b]]
        key = list(zip(r[:, 5].tolist(), (-r[:, 4]).tolist()))
        assert key == sorted(key)


def test_nms_ties_all_equal_scores():
    from oracle import detector_oracle as O
    from transparent_object_detection_b200 import DecodeBox
    rng = np.random.default_rng(5)
    A, nc = 2000, 3
    pred = np.zeros((1, A, 4 + nc), np.float32)
    pred[0, :, 0:2] = rng.uniform(0.2, 0.8, (A, 2)).astype(np.float32)
    pred[0, :, 2:4] = rng.uniform(0.05, 0.2, (A, 2)).astype(np.float32)
    pred[0, np.arange(A), 4 + rng.integers(0, nc, A)] = 0.5
    pred[0, ::7, :4] = pred[0, 0, :4]                                   # duplicated boxes
    want = O.nms_keep_indices(pred, nc, 0.5, 0.5)
    keep_idx, keep_count, _ = DecodeBox(nc, (640, 640)).nms_device(torch.from_numpy(pred.copy()).cuda(), nc, 0.5, 0.5)
    n = int(keep_count[0])
    assert n == len(want[0]) and np.array_equal(keep_idx[0, :n].cpu().numpy(), want[0])


# ------------------------------------------------------------------------------------------ whole network
def _rel_stats(got, want):
    err = np.abs(got - want)
    return err.max() / np.abs(want).max(), np.sqrt((err ** 2).mean()) / np.sqrt((want ** 2).mean())


def test_network_scale_n_matches_reference_fixture(golden):
    from oracle import synth
    from transparent_object_detection_b200 import BaseModel
    g = golden("net_n_96x128.npz")
    C_, d, m = synth.SCALES["n"]
    model = BaseModel(80, C_, d, m)
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.make_state_dict(80, C_, d, m, seed=0).items()})
    model.eval()
    x = torch.from_numpy(synth.make_images(2, 96, 128, seed=2)).cuda()
    out = model(x)
    eng = model.engine(2, 96, 128, x.device)
    torch.cuda.synchronize()
    for name in ("p3", "p4", "p5", "h2", "h4", "h6"):
        mx, rms = _rel_stats(eng.feature_nchw(name).cpu().numpy(), g[name])
        print(f"scale n {name}: max-abs err / abs-max {mx:.4f}, RMS-rel {rms:.4f}")
        assert mx <= 0.02 and rms <= 0.01, (name, mx, rms)
    raw = [r.float().cpu().numpy() for r in eng.raw_maps_nchw()]
    for i in range(3):
        assert np.abs(raw[i] - g[f"raw{i}"]).max() <= 0.08, (i, np.abs(raw[i] - g[f"raw{i}"]).max())
    o = out.cpu().numpy()
    assert o.shape == g["out"].shape
    print(f"scale n out: box err {np.abs(o[:, :4] - g['out'][:, :4]).max():.4f} px, score err {np.abs(o[:, 4:] - g['out'][:, 4:]).max():.2e}")
    assert np.abs(o[:, :4] - g["out"][:, :4]).max() <= 1.0, np.abs(o[:, :4] - g["out"][:, :4]).max()
    assert np.abs(o[:, 4:] - g["out"][:, 4:]).max() <= 5e-3
    # train mode returns the raw maps (model/head.py:50-51)
    model.train()
    tr = model(x)
    assert [tuple(t.shape) for t in tr] == [(2, 144, 12, 16), (2, 144, 6, 8), (2, 144, 3, 4)]


def test_network_scale_s_640_config1(golden):
    """BASELINE config 1: scale s, 1x3x640x640; features/boxes/scores within tolerance of the CPU oracle, and
    the full pipeline's detections consistent with the oracle's NMS fed OUR decoded tensor (bit-exact)."""
    from oracle import detector_oracle as O, synth
    from transparent_object_detection_b200 import BaseModel, DecodeBox
    g = golden("config1_s_640.npz")
    C_, d, m = synth.SCALES["s"]
    sd = synth.make_state_dict(80, C_, d, m, seed=0)
    model = BaseModel(80, C_, d, m).eval()
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    x = torch.from_numpy(synth.make_images(1, 640, 640, seed=2))
    out = model(x.cuda())
    o = out.cpu().numpy()
    assert o.shape == (1, 84, 8400)
    print(f"config 1: box err {np.abs(o[:, :4, ::16] - g['out_sub'][:, :4]).max():.4f} px, "
          f"score err {np.abs(o[:, 4:, ::16] - g['out_sub'][:, 4:]).max():.2e}")
    assert np.abs(o[:, :4, ::16] - g["out_sub"][:, :4]).max() <= 1.0
    assert np.abs(o[:, 4:, ::16] - g["out_sub"][:, 4:]).max() <= 5e-3
    db = DecodeBox(80, (640, 640))
    dec = db.decode_box(out)
    dec_np = dec.cpu().numpy()
    want = O.non_max_suppression(dec_np.copy(), 80, (640, 640), (480, 640), True, 0.05, 0.5)
    got = db.non_max_suppression(dec, 80, (640, 640), np.array((480, 640)), True, conf_thres=0.05, nms_thres=0.5)
    assert_dets_equal(got, want)
    # against the reference's own rows for this image (fixture): every reference row has a same-class row of ours that it
    # overlaps (a bf16 score flip swaps two overlapping survivors); the strict IoU >= 0.9 agreement is printed
    ref_rows = g["nms_cb_rows"][:int(g["nms_cb_counts"][0])]
    strict, loose = _match_rate(got[0], ref_rows, 0.9), _match_rate(got[0], ref_rows, 0.5)
    print(f"config 1 rows: ours {got[0].shape[0]}, reference {len(ref_rows)}, matched at IoU >= 0.9: {strict:.3f}, at IoU >= 0.5: {loose:.3f}")
    assert got[0] is not None and loose >= 0.95 and strict >= 0.75 and abs(got[0].shape[0] - len(ref_rows)) <= 0.1 * len(ref_rows) + 5


def test_detector_graph_matches_eager_and_oracle_nms():
    """Detector.detect (one CUDA graph: network + fused decode + NMS) == eager API chain, batch 3."""
    from oracle import detector_oracle as O, synth
    from transparent_object_detection_b200 import BaseModel, DecodeBox, Detector
    C_, d, m = synth.SCALES["n"]
    model = BaseModel(80, C_, d, m).eval()
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.make_state_dict(80, C_, d, m, seed=0).items()})
    x = torch.from_numpy(synth.make_images(3, 160, 192, seed=9))
    det = Detector(model, (160, 192), confidence=0.01, nms_iou=0.5, letterbox_image=True)
    got = det.detect(x.pin_memory(), image_shape=(300, 400))
    got2 = det.detect(x.cuda(), image_shape=(300, 400))          # replay is deterministic
    db = DecodeBox(80, (160, 192))
    dec = db.decode_box(model(x.cuda()))
    want = O.non_max_suppression(dec.cpu().numpy(), 80, (160, 192), (300, 400), True, 0.01, 0.5)
    assert_dets_equal(got, want)
    assert_dets_equal(got2, want)
    assert any(w is not None for w in want)


@pytest.mark.parametrize("scale", ["n", "s"])
def test_fused_head_decode_equals_conv_then_decode_bit_for_bit(scale):
    """tod_conv2d_head_decode (last conv of a head tower + its share of the decode in the conv epilogue) and
    tod_conv2d_tail1x1_box_decode (the box tower's last TWO convs + decode in one kernel) must give exactly the
    candidates of the unfused path (f32 raw maps -> tod_head_decode), at a batch / size whose flat 128-row tiles straddle
    images and end in a partial tile."""
    from oracle import synth
    from transparent_object_detection_b200 import BaseModel
    C_, d, m = synth.SCALES[scale]
    model = BaseModel(80, C_, d, m).eval()
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.make_state_dict(80, C_, d, m, seed=0).items()})
    eng = model.engine(3, 96, 160)
    x = torch.from_numpy(synth.make_images_u8(3, 96, 160, seed=5)).cuda()
    eng.run_network(x)
    eng.run_decode(False, False, True)
    torch.cuda.synchronize()
    want = [t.clone() for t in (eng.cand_box, eng.cand_conf, eng.cand_cls)]
    for t in (eng.cand_box, eng.cand_conf, eng.cand_cls):
        t.fill_(-7)
    eng.run_network(x, fused_decode=True)
    torch.cuda.synchronize()
    assert len(eng.tail_box) == 3                       # the box towers went through the fused tail
    assert torch.equal(eng.cand_cls, want[2])
    assert torch.equal(eng.cand_conf, want[1])
    assert torch.equal(eng.cand_box, want[0])
    assert float(want[1].max()) > 0 and int(want[2].max()) > 0


# ------------------------------------------------------------------------------------------ uint8 input path
@pytest.mark.parametrize("B,H,W,cout", [(2, 64, 96, 32), (1, 32, 32, 16), (1, 640, 640, 32), (3, 38, 50, 64), (1, 64, 64, 128)])
def test_stem_u8_tensor_core(B, H, W, cout):
    """uint8 NHWC stem on tcgen05 == conv2d of (u8 / 255) with the SAME bf16-rounded (weights / 255): the only other
    differences are fp32 accumulation order, the tanh-form SiLU (<= |x| * 2.5e-4) and bf16 output rounding."""
    _lib, L = _engine_parts()
    g = torch.Generator().manual_seed(3)
    u8 = torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8).cuda()
    w = torch.randn((cout, 3, 3, 3), generator=g) * (2.0 / 27) ** 0.5
    b = torch.randn((cout,), generator=g) * 0.3
    out = torch.zeros((B, H // 2, W // 2, cout), dtype=torch.bfloat16).cuda()
    wh = w.reshape(cout, 27).contiguous()
    _lib.check(L.tod_stem_conv_nhwc_u8(u8.data_ptr(), wh.data_ptr(), b.data_ptr(), out.data_ptr(), B, H, W, cout, cout,
                                       torch.cuda.current_stream().cuda_stream), "stem_u8")
    wq = (w * np.float32(1.0 / 255.0)).to(torch.bfloat16).float().cuda()
    x = u8.permute(0, 3, 1, 2).float()
    want = F.silu(F.conv2d(x, wq, b.cuda(), stride=2, padding=1)).permute(0, 2, 3, 1)
    err = (out.float() - want).abs()
    assert float((err - 1e-2 * want.abs()).max()) <= 1e-2, float(err.max())
    # and against the fp32 reference arithmetic (x / 255 with unrounded weights): bf16-level agreement
    ref = F.silu(F.conv2d(x / 255.0, w.cuda(), b.cuda(), stride=2, padding=1)).permute(0, 2, 3, 1)
    assert float(((out.float() - ref).abs() - 2e-2 * ref.abs()).max()) <= 2e-2


def test_detector_uint8_pipeline_matches_float_path_and_is_order_safe():
    """uint8 (B, H, W, 3) input == float32 (u8 / 255) input within the stated bf16 tolerance at the network output;
    two batches submitted back to back (double-buffered upload) give exactly the rows of two synchronous calls."""
    from oracle import synth
    from transparent_object_detection_b200 import BaseModel, Detector
    C_, d, m = synth.SCALES["n"]
    model = BaseModel(80, C_, d, m).eval()
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.make_state_dict(80, C_, d, m, seed=0).items()})
    g = torch.Generator().manual_seed(5)
    u8a = torch.randint(0, 256, (3, 160, 192, 3), generator=g, dtype=torch.uint8)
    u8b = torch.randint(0, 256, (3, 160, 192, 3), generator=g, dtype=torch.uint8)
    det = Detector(model, (160, 192), confidence=0.01, nms_iou=0.5, letterbox_image=True)
    ra, rb = det.detect(u8a.pin_memory()), det.detect(u8b.pin_memory())
    pa = det.submit(u8a.pin_memory())
    pb = det.submit(u8b.pin_memory())
    assert_dets_equal(det.collect(pa), ra)
    assert_dets_equal(det.collect(pb), rb)
    assert any(r is not None for r in ra)
    # network outputs of the two input formats agree to the bf16 tolerance
    eng = model.engine(3, 160, 192)
    eng.run_network(u8a.cuda()); eng.run_decode(True, False, False)
    torch.cuda.synchronize()
    o_u8 = eng.head_out.clone()
    xf = (u8a.permute(0, 3, 1, 2).float() / 255.0).contiguous().cuda()
    eng.run_network(xf); eng.run_decode(True, False, False)
    torch.cuda.synchronize()
    o_f = eng.head_out
    assert float((o_u8[:, :4] - o_f[:, :4]).abs().max()) <= 1.0
    assert float((o_u8[:, 4:] - o_f[:, 4:]).abs().max()) <= 5e-3


# ------------------------------------------------------------------------------------------------ f3: device correct_boxes
@pytest.mark.parametrize("letterbox", [True, False])
def test_correct_boxes_device_bit_exact_vs_numpy_flow(letterbox):
    """tod_correct_boxes against the numpy dtype flow of utils/bbox_utils.py:84-117,176-180 (host port pinned to the
    reference fixture in tests/test_host_cpu.py): random kept rows, one (h, w) per image, bit-exact float32 rows."""
    import ctypes as C
    from transparent_object_detection_b200 import DecodeBox, lib
    from transparent_object_detection_b200._lib import check
    from transparent_object_detection_b200.model import box_correction_params
    g = torch.Generator().manual_seed(11)
    B, A = 5, 300
    xy = torch.rand((B, A, 2), generator=g)
    wh = torch.rand((B, A, 2), generator=g) * 0.5
    dets = torch.cat([xy - wh / 2, xy + wh / 2, torch.rand((B, A, 1), generator=g),
                      torch.randint(0, 80, (B, A, 1), generator=g).float()], 2).contiguous()
    counts = torch.tensor([300, 0, 17, 1, 250], dtype=torch.int32)
    shapes = np.array([[480, 640], [640, 480], [1080, 1920], [333, 777], [640, 640]])
    input_shape = (640, 640)
    prm = torch.from_numpy(box_correction_params(input_shape, shapes, B, letterbox)).cuda()
    d_dev, c_dev = dets.cuda(), counts.cuda()
    out = torch.full_like(d_dev, -7.0)
    check(lib().tod_correct_boxes(d_dev.data_ptr(), c_dev.data_ptr(), B, A, prm.data_ptr(), int(letterbox), out.data_ptr(),
                                  torch.cuda.current_stream().cuda_stream), "tod_correct_boxes")
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    host = dets.numpy().copy()
    for i in range(B):
        n = int(counts[i])
        rows = host[i, :n].copy()
        if n:
            box_xy, box_wh = (rows[:, 0:2] + rows[:, 2:4]) / 2, rows[:, 2:4] - rows[:, 0:2]
            rows[:, :4] = DecodeBox.correct_boxes(box_xy, box_wh, input_shape, shapes[i], letterbox)
        assert np.array_equal(got[i, :n], rows), i
        assert np.all(got[i, n:] == -7.0)                      # rows past the count are not touched


def test_detector_device_corrected_rows_equal_host_corrected_rows():
    import transparent_object_detection_b200 as T
    from oracle import synth
    C_, d, m = synth.SCALES["n"]
    model = T.BaseModel(80, C_, d, m).eval()
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.make_state_dict(80, C_, d, m, seed=0).items()})
    x = torch.from_numpy(synth.make_images_u8(3, 64, 96, seed=9))
    shapes = np.array([[48, 96], [200, 150], [64, 96]])
    for lb in (True, False):
        det = T.Detector(model, (64, 96), confidence=0.01, nms_iou=0.5, letterbox_image=lb)
        dev_rows = det.detect(x, shapes)                        # shapes known at submit: tod_correct_boxes
        host_rows = det.collect(det.submit(x), shapes)          # shapes given at collect: numpy on the host
        assert sum(r is not None for r in host_rows) > 0
        for a, b in zip(dev_rows, host_rows):
            assert (a is None) == (b is None)
            if a is not None:
                assert np.array_equal(a, b)


# ------------------------------------------------------------------------------------------ BASELINE configs 3 and 4
def test_sharded_image_ranges_equal_the_single_batch_result():
    """BASELINE config 3 on one GPU: the rank partition of a batch (shard_range, W = 2 and 4), each shard run as its own
    batch, concatenated in rank order == the one-batch result, row for row (every op is per-image)."""
    import transparent_object_detection_b200 as T
    from oracle import synth
    C_, d, m = synth.SCALES["n"]
    model = T.BaseModel(80, C_, d, m).eval()
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.make_state_dict(80, C_, d, m, seed=0).items()})
    x = torch.from_numpy(synth.make_images_u8(12, 160, 160, seed=4))
    det = T.Detector(model, (160, 160), confidence=0.01, nms_iou=0.5)
    whole = det.detect(x)
    assert sum(r is not None for r in whole) > 0
    for world in (2, 4):
        parts = []
        for rank in range(world):
            lo, hi = T.shard_range(12, rank, world)
            parts.extend(det.detect(x[lo:hi].contiguous()))
        assert_dets_equal(parts, whole)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_second_device_in_the_same_process_gives_the_same_detections():
    """One process driving two GPUs (BaseModel.engine(device=...), Detector on tensors of either device): the one-time kernel
    attribute setup (dynamic shared memory limits of the tcgen05 / pooling / NMS kernels, CTA-pair instantiations included) is
    per device, so the second device must run the same plan and return the same rows bit for bit."""
    import transparent_object_detection_b200 as T
    from oracle import synth
    C_, d, m = synth.SCALES["n"]
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in synth.make_state_dict(80, C_, d, m, seed=0).items()}
    x = torch.from_numpy(synth.make_images_u8(4, 160, 160, seed=12))
    rows = []
    for dev in (0, 1):
        with torch.cuda.device(dev):
            model = T.BaseModel(80, C_, d, m).eval()
            model.load_state_dict(sd)
            det = T.Detector(model, (160, 160), confidence=0.01, nms_iou=0.5)
            rows.append(det.detect(x.to(f"cuda:{dev}")))
            head = model(torch.from_numpy(synth.images_u8_to_f32(x.numpy())).to(f"cuda:{dev}"))
            rows.append([head.cpu().numpy()])
    assert sum(r is not None for r in rows[0]) > 0
    assert_dets_equal(rows[2], rows[0])
    assert np.array_equal(rows[3][0], rows[1][0])


def test_upstream_style_detect_image_and_get_fps(tmp_path):
    """The facade predict.py is written against (predict.py:105,130,156,168): detect_image(image, crop, count) returns
    the annotated PIL image -- same size, pixels changed exactly where boxes were kept, crops written -- and get_FPS(image, n)
    a positive time per image; the get_map form of detect_image keeps working through the same method."""
    import transparent_object_detection_b200 as T
    from PIL import Image
    from oracle import synth
    C_, d, m = synth.SCALES["n"]
    model = T.BaseModel(80, C_, d, m).eval()
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.make_state_dict(80, C_, d, m, seed=0).items()})
    det = T.Detector(model, (160, 224), confidence=0.01, nms_iou=0.5, max_boxes=20)
    rng = np.random.default_rng(3)
    img = Image.fromarray(rng.integers(0, 256, (120, 200, 3), dtype=np.uint8))
    rows = det.top_boxes(det.detect_image_rows(img))
    assert rows is not None and 0 < len(rows) <= 20
    out = det.annotate_image(img.copy(), crop=True, count=True, crop_dir=str(tmp_path / "crops"))
    assert isinstance(out, Image.Image) and out.size == img.size
    assert np.any(np.asarray(out) != np.asarray(img))
    assert len(list((tmp_path / "crops").iterdir())) > 0
    out2 = det.detect_image(img.copy())                      # upstream call form, positional image only
    assert isinstance(out2, Image.Image) and np.array_equal(np.asarray(out2), np.asarray(det.annotate_image(img.copy())))
    res = det.detect_image(7, img, [], {i: i + 1 for i in range(80)})          # get_map.py form
    assert isinstance(res, list) and len(res) > 0 and res[0]["image_id"] == 7 and set(res[0]) == {"image_id", "category_id", "bbox", "score"}
    t = det.get_FPS(img, 3)
    assert 0.0 < t < 5.0
    blank = det.__class__(model, (160, 224), confidence=0.999, nms_iou=0.5)   # nothing passes: the image comes back untouched
    assert np.array_equal(np.asarray(blank.detect_image(img.copy())), np.asarray(img))


@pytest.mark.parametrize("case,reg_max", [("net", 16), ("syn", 16), ("nodfl", 1)])
@pytest.mark.parametrize("on_cpu", [False, True], ids=["cuda_in", "cpu_in"])
def test_loss_bbox_decode_vs_reference_fixture(golden, case, reg_max, on_cpu):
    """LossDecode.bbox_decode (tod_loss_bbox_decode) against the reference's own Loss.bbox_decode
    (model/loss.py:333-337; fixture from oracle/make_golden_loss.py).  float32, stated tolerance 2e-5 grid units absolute
    (the reference's `.softmax(3).matmul(proj)` sums its 16 terms in ATen's order; the kernel sums them in index order)."""
    import transparent_object_detection_b200 as T
    g = golden("loss_bbox_decode.npz")
    ap, pd, want = (torch.from_numpy(g[f"{case}_{k}"]) for k in ("anchor_points", "pred_dist", "boxes"))
    head = type("H", (), {"ch": reg_max, "nc": 80, "stride": [8, 16, 32]})()
    ld = T.LossDecode(type("M", (), {"head": head})())
    assert ld.reg_max == reg_max and ld.use_dfl == (reg_max > 1)
    got = ld.bbox_decode(ap if on_cpu else ap.cuda(), pd if on_cpu else pd.cuda())
    assert got.is_cuda != on_cpu and tuple(got.shape) == tuple(want.shape)
    err = float((got.cpu() - want).abs().max())
    assert err <= (2e-5 if reg_max > 1 else 0.0), err
    with pytest.raises(ValueError):
        ld.bbox_decode(ap, pd[:, :, :-1])


def test_sm_budget_changes_the_grids_not_the_results():
    """tod_set_sm_budget(n): the persistent kernels (stem, convs) size their grids by n SMs -- several batches in flight can
    then share the device side by side.  Which CTA computes a tile never changes what is computed: detections must be
    bit-identical for every budget (and the budget is read at capture time, so each Detector captures its own graphs)."""
    import transparent_object_detection_b200 as T
    from oracle import synth
    L = T.lib()
    C_, d, m = synth.SCALES["n"]
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in synth.make_state_dict(80, C_, d, m, seed=0).items()}
    x = torch.from_numpy(synth.make_images_u8(6, 160, 224, seed=9))
    try:
        results = []
        for budget in (0, 37, 5, 1):
            assert L.tod_set_sm_budget(budget) == 0 and L.tod_get_sm_budget() == budget
            model = T.BaseModel(80, C_, d, m).eval()          # a fresh plan: engines and their graphs are cached per model
            model.load_state_dict(sd)
            det = T.Detector(model, (160, 224), confidence=0.01, nms_iou=0.5)
            results.append(det.detect(x))
        assert sum(r is not None for r in results[0]) > 0
        for r in results[1:]:
            assert_dets_equal(r, results[0])
        assert L.tod_set_sm_budget(-1) != 0          # rejected, the budget stays
        assert L.tod_get_sm_budget() == 1
    finally:
        L.tod_set_sm_budget(0)


def test_network_1280_config4_against_live_oracle():
    """BASELINE config 4 geometry (1280x1280: A = 33600, SPPF planes 40x40) at scale n, one image, against the CPU
    oracle evaluated in the test: boxes <= 1 px, scores <= 5e-3; NMS on OUR decoded tensor bit-exact with the oracle's."""
    from oracle import detector_oracle as O, synth
    from transparent_object_detection_b200 import BaseModel, DecodeBox
    C_, d, m = synth.SCALES["n"]
    sd = synth.make_state_dict(80, C_, d, m, seed=0)
    model = BaseModel(80, C_, d, m).eval()
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    x = torch.from_numpy(synth.make_images(1, 1280, 1280, seed=5))
    out = model(x.cuda())
    with torch.no_grad():
        want = O.forward(sd, x, 80, d)
    o = out.cpu()
    assert tuple(o.shape) == (1, 84, 33600)
    assert float((o[:, :4] - want[:, :4]).abs().max()) <= 1.0
    assert float((o[:, 4:] - want[:, 4:]).abs().max()) <= 5e-3
    db = DecodeBox(80, (1280, 1280))
    dec = db.decode_box(out)
    want_rows = O.non_max_suppression(dec.cpu().numpy().copy(), 80, (1280, 1280), (720, 1280), True, 0.01, 0.5)
    got_rows = db.non_max_suppression(dec, 80, (1280, 1280), np.array((720, 1280)), True, conf_thres=0.01, nms_thres=0.5)
    assert_dets_equal(got_rows, want_rows)


@pytest.mark.parametrize("scale", ["n", "s", "m", "l", "x"])
def test_every_scale_against_live_oracle(scale):
    """All five scales of config.yaml (n (16,1,1.0) ... x (96,3,0.5)): depth 1-3, widths that are not powers
    of two (48, 96, 576 channels), against the CPU oracle evaluated here.  n and s meet the stated absolute tolerance
    (features <= 2 % of abs-max, boxes <= 1 px, scores <= 5e-3).  The deeper random-init networks amplify any perturbation (activations reach
    |x| ~ 100 at scale l), so that bf16 STORAGE alone -- the fp32 oracle re-evaluated with the build's rounding points,
    oracle.bf16_emulation -- already deviates from fp32 by up to 15 px / 0.11 in score there; for those the kernel must
    stay within 2.5x that inherent deviation on the raw head maps and on every stage feature (or 4 % of abs-max)."""
    from oracle import detector_oracle as O, synth
    from transparent_object_detection_b200 import BaseModel
    C_, d, m = synth.SCALES[scale]
    sd = synth.make_state_dict(80, C_, d, m, seed=0)
    model = BaseModel(80, C_, d, m).eval()
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    x = torch.from_numpy(synth.make_images(2, 64, 96, seed=7))
    out = model(x.cuda()).cpu()
    eng = model.engine(2, 64, 96, torch.device("cuda", torch.cuda.current_device()))
    with torch.no_grad():
        feats = O.backbone(sd, x, d)
        necks = O.neck(sd, feats, d)
        raw = O.head_raw(sd, necks)
        want = O.head_decode(raw, 80)
        with O.bf16_emulation():
            feats_emu = O.backbone(sd, x, d)
            necks_emu = O.neck(sd, feats_emu, d)
            raw_emu = O.head_raw(sd, necks_emu)
    assert tuple(out.shape) == tuple(want.shape)
    for name, ref, emu in zip(("p3", "p4", "p5", "h2", "h4", "h6"), list(feats) + list(necks), list(feats_emu) + list(necks_emu)):
        err = float((eng.feature_nchw(name).cpu() - ref).abs().max())
        inherent = float((emu - ref).abs().max())                  # what bf16 storage alone costs at this stage
        lim = 0.02 * float(ref.abs().max()) if scale in ("n", "s") else max(0.04 * float(ref.abs().max()), 2.5 * inherent)
        assert err <= lim, (name, err, inherent, float(ref.abs().max()))
    for i, r in enumerate(eng.raw_maps_nchw()):
        ours = float((r.float().cpu() - raw[i]).abs().max())
        inherent = float((raw_emu[i] - raw[i]).abs().max())
        assert ours <= 2.5 * inherent + 0.02, (i, ours, inherent)
    if scale in ("n", "s"):
        assert float((out[:, :4] - want[:, :4]).abs().max()) <= 1.0
        assert float((out[:, 4:] - want[:, 4:]).abs().max()) <= 5e-3


@pytest.mark.parametrize("nc", [1, 3, 20])
def test_network_with_few_classes_against_live_oracle(nc):
    """Class counts other than 80 (the class conv is padded to 16 output channels, the class tower is max(C3, nc) wide):
    head tensor within tolerance of the CPU oracle, Detector rows == the oracle's NMS on OUR decoded tensor, bit-exact."""
    from oracle import detector_oracle as O, synth
    from transparent_object_detection_b200 import BaseModel, DecodeBox, Detector
    C_, d, m = synth.SCALES["n"]
    sd = synth.make_state_dict(nc, C_, d, m, seed=0)
    model = BaseModel(nc, C_, d, m).eval()
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    x = torch.from_numpy(synth.make_images(2, 96, 128, seed=3))
    out = model(x.cuda())
    with torch.no_grad():
        want = O.forward(sd, x, nc, d)
    o = out.cpu()
    assert tuple(o.shape) == (2, 4 + nc, 12 * 16 + 6 * 8 + 3 * 4)
    assert float((o[:, :4] - want[:, :4]).abs().max()) <= 1.0
    assert float((o[:, 4:] - want[:, 4:]).abs().max()) <= 5e-3
    db = DecodeBox(nc, (96, 128))
    dec = db.decode_box(out)
    conf = float(dec[:, :, 4:].max()) * 0.5                    # a threshold that keeps some anchors whatever the init
    want_rows = O.non_max_suppression(dec.cpu().numpy().copy(), nc, (96, 128), (96, 128), True, conf, 0.5)
    det = Detector(model, (96, 128), confidence=conf, nms_iou=0.5)
    assert_dets_equal(det.detect(x), want_rows)
    assert any(r is not None for r in want_rows)


# ------------------------------------------------------------------------------------------------ f3: packing / top-k
@pytest.mark.parametrize("B,A,max_boxes", [(3, 64, 0), (5, 300, 100), (2, 8400, 100), (2, 8400, 300), (1, 20000, 50), (4, 33, 1000)])
def test_pack_detections_and_topk_bit_exact(B, A, max_boxes):
    """tod_pack_detections: compaction in NMS order (max_boxes 0) / top-k by score descending, ties in kept order
    (utils/callbacks.py:159-166 with a DEFINED tie order), offsets = exclusive scan of the per-image counts."""
    _lib, L = _engine_parts()
    g = np.random.Generator(np.random.PCG64(B * 1000 + A))
    rows = g.random((B, A, 6), dtype=np.float32)
    rows[..., 4] = np.round(rows[..., 4] * 40) / 40                 # many exact score ties
    counts = g.integers(0, A + 1, size=B).astype(np.int32)
    counts[0] = A
    if B > 1:
        counts[1] = 0
    d_rows, d_cnt = torch.from_numpy(rows).cuda(), torch.from_numpy(counts).cuda()
    d_off = torch.full((B + 1,), -7, dtype=torch.int32, device="cuda")
    d_out = torch.full((B * A, 6), -1.0, dtype=torch.float32, device="cuda")
    wb = int(L.tod_pack_workspace_bytes(B, A))
    work = torch.zeros(max(wb, 8), dtype=torch.uint8, device="cuda")
    _lib.check(L.tod_pack_detections(d_rows.data_ptr(), d_cnt.data_ptr(), B, A, max_boxes, d_off.data_ptr(), d_out.data_ptr(),
                                     work.data_ptr(), work.numel(), torch.cuda.current_stream().cuda_stream), "pack")
    torch.cuda.synchronize()
    off, out = d_off.cpu().numpy(), d_out.cpu().numpy()
    want_rows = []
    for b in range(B):
        r = rows[b, :counts[b]]
        if max_boxes > 0:
            order = np.lexsort((np.arange(len(r)), -r[:, 4].astype(np.float64)))[:max_boxes]     # score desc, then kept order
            r = r[order]
        want_rows.append(r)
    want_off = np.concatenate(([0], np.cumsum([len(r) for r in want_rows]))).astype(np.int32)
    assert np.array_equal(off, want_off)
    assert np.array_equal(out[:want_off[-1]], np.concatenate(want_rows) if want_off[-1] else np.zeros((0, 6), np.float32))
    assert (out[want_off[-1]:] == -1.0).all()                        # nothing written past the packed rows


def _small_detector(depth=3, conf=0.01):
    from oracle import synth
    from transparent_object_detection_b200 import BaseModel, Detector
    C_, d, m = synth.SCALES["n"]
    model = BaseModel(80, C_, d, m).eval()
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.make_state_dict(80, C_, d, m, seed=0).items()})
    return model, Detector(model, (160, 192), confidence=conf, nms_iou=0.5, letterbox_image=True, pipeline_depth=depth)


def test_detector_topk_on_device_matches_reference_host_expression():
    """Detector.detect(max_boxes=k) / detect_top (device top-k inside the captured graph) against the reference's own host
    expression on the full rows (Detector.top_boxes = np.argsort(conf)[::-1][:k]): same rows as SETS, score descending, and
    identical arrays wherever the scores are distinct (numpy's unstable argsort defines no tie order)."""
    model, det = _small_detector()
    g = torch.Generator().manual_seed(9)
    u8 = torch.randint(0, 256, (3, 160, 192, 3), generator=g, dtype=torch.uint8).pin_memory()
    full = det.detect(u8, (160, 192))
    assert max(len(r) for r in full if r is not None) > 12
    for k in (5, 12, 1000):
        det.max_boxes = k
        top = det.detect(u8, (160, 192), max_boxes=k)
        trip = det.detect_top(u8, (160, 192))
        for rows, t, tr in zip(full, top, trip):
            if rows is None:
                assert t is None and tr is None
                continue
            want = det.top_boxes(rows)
            assert t.shape == want.shape and np.all(np.diff(t[:, 4]) <= 0)
            assert np.array_equal(np.sort(t[:, 4]), np.sort(want[:, 4]))
            if len(np.unique(rows[:, 4])) == len(rows):
                assert np.array_equal(t, want)
            assert np.array_equal(tr[0], t[:, 5].astype("int32")) and np.array_equal(tr[1], t[:, 4]) and np.array_equal(tr[2], t[:, :4])


def test_detector_more_submits_than_pipeline_depth_never_mixes_batches():
    """a plan (arena + result buffers) is reused every `pipeline_depth` submits; a handle that is still
    uncollected then must keep ITS rows (they are fetched into the handle before the plan is reused)."""
    model, det = _small_detector(depth=2)
    g = torch.Generator().manual_seed(10)
    batches = [torch.randint(0, 256, (2, 160, 192, 3), generator=g, dtype=torch.uint8).pin_memory() for _ in range(5)]
    want = [det.detect(b, (160, 192)) for b in batches]
    assert any(r is not None for w in want for r in w)
    pend = [det.submit(b, (160, 192)) for b in batches]          # 5 submits, depth 2, nothing collected yet
    for p, w in reversed(list(zip(pend, want))):                 # collected in reverse order
        assert_dets_equal(det.collect(p), w)
        assert_dets_equal(det.collect(p), w)                     # idempotent
    # thresholds changed with batches in flight: the re-capture must not disturb them
    p0 = det.submit(batches[0], (160, 192))
    det.confidence = 0.02
    p1 = det.submit(batches[1], (160, 192))
    assert_dets_equal(det.collect(p0), want[0])
    r1 = det.collect(p1)
    for a, b in zip(r1, want[1]):
        if b is not None:
            keep = b[b[:, 4] >= np.float32(0.02)]
            assert (a is None and len(keep) == 0) or np.array_equal(a, keep)


# ------------------------------------------------------------------------------------------------ parity at the benchmarked shapes
def _match_rate(got, want, iou_thr=0.9):
    """Fraction of the oracle's rows that have a same-class row of ours with IoU >= iou_thr (rows [y1, x1, y2, x2, conf, cls])."""
    if want is None or len(want) == 0:
        return 1.0
    if got is None:
        return 0.0
    hit = 0
    for w in want:
        c = got[got[:, 5] == w[5]]
        if len(c) == 0:
            continue
        iy = np.clip(np.minimum(c[:, 2], w[2]) - np.maximum(c[:, 0], w[0]), 0, None)
        ix = np.clip(np.minimum(c[:, 3], w[3]) - np.maximum(c[:, 1], w[1]), 0, None)
        inter = iy * ix
        union = (c[:, 2] - c[:, 0]) * (c[:, 3] - c[:, 1]) + (w[2] - w[0]) * (w[3] - w[1]) - inter
        hit += bool((inter / np.maximum(union, 1e-9)).max() >= iou_thr)
    return hit / len(want)


def _benchmark_shape_parity(B, size, depth, sample, conf=0.05, iou=0.5):
    """The EXACT bench path -- Detector.submit / collect with `depth` batches in flight on the uint8 batch (graph: u8 stem,
    forked towers, fused decode, NMS, device un-letterbox, packed rows) -- at the benchmarked shape, against the fp32 CPU
    oracle on `sample` images.  Stated tolerance: boxes <= 1 px, scores <= 5e-3, features see the network tests."""
    from oracle import detector_oracle as O, synth
    from transparent_object_detection_b200 import BaseModel, DecodeBox, Detector
    C_, d, m = synth.SCALES["s"]
    sd = synth.make_state_dict(80, C_, d, m, seed=0)
    model = BaseModel(80, C_, d, m).eval()
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    det = Detector(model, (size, size), confidence=conf, nms_iou=iou, letterbox_image=True, pipeline_depth=depth)
    hosts = [torch.from_numpy(synth.make_images_u8(B, size, size, seed=3 + j)).pin_memory() for j in range(2)]
    pend = [det.submit(hosts[i & 1], (size, size)) for i in range(depth + 2)]          # more than `depth`: plans are reused
    results = [det.collect(p) for p in pend]
    for i, r in enumerate(results):                                                   # same input -> same rows, every plan instance
        assert_dets_equal(r, results[i & 1])
    # (1) the graph's rows == NMS of OUR OWN Head tensor through the eager chain (unfused decode kernel), bit for bit
    eng = model.engine(B, size, size, instance=1)
    x = eng.input_buffer("u8", 0)
    x.copy_(hosts[0])
    eng.run_network(x)
    eng.run_decode(True, False, False)
    torch.cuda.synchronize()
    head = eng.head_out.clone()
    db = DecodeBox(80, (size, size))
    chain = db.non_max_suppression(db.decode_box(head[sample]), 80, (size, size), np.array((size, size)), True, conf, iou)
    assert_dets_equal([results[0][i] for i in sample], chain)
    # (2) Head tensor and rows against the fp32 oracle on the sampled images
    xs = torch.from_numpy(synth.images_u8_to_f32(hosts[0].numpy()[sample]))
    sdt = {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}
    with torch.no_grad():
        want = O.forward(sdt, xs, 80, d)
    got = head[sample].cpu()
    box_err, score_err = float((got[:, :4] - want[:, :4]).abs().max()), float((got[:, 4:] - want[:, 4:]).abs().max())
    want_rows = O.non_max_suppression(O.decode_box(want, (size, size)).numpy(), 80, (size, size), (size, size), True, conf, iou)
    strict = [_match_rate(results[0][i], w, 0.9) for i, w in zip(sample, want_rows)]
    loose = [_match_rate(results[0][i], w, 0.5) for i, w in zip(sample, want_rows)]
    kept = [(0 if results[0][i] is None else len(results[0][i]), 0 if w is None else len(w)) for i, w in zip(sample, want_rows)]
    print(f"benchmark-shape parity B={B} {size}x{size}: box err {box_err:.3f} px, score err {score_err:.2e}, "
          f"kept (ours, oracle) {kept}; the oracle's rows with a same-class row of ours at IoU >= 0.9: {np.mean(strict):.3f}, "
          f"at IoU >= 0.5 (a bf16 score flip swaps two overlapping survivors): {np.mean(loose):.3f}")
    assert box_err <= 1.0 and score_err <= 5e-3, (box_err, score_err)
    assert np.mean(loose) >= 0.95 and np.mean(strict) >= 0.75, (strict, loose)


def test_config2_benchmark_path_parity():
    """BASELINE config 2: scale s, batch 64, 640x640, pipeline depth 4 (bench.py's e2e leg), 8 sampled images."""
    _benchmark_shape_parity(64, 640, 4, [0, 7, 13, 22, 31, 40, 55, 63])


def test_config4_benchmark_path_parity():
    """BASELINE config 4: scale s, batch 16, 1280x1280 (SPPF planes 40x40, A = 33600), 2 sampled images."""
    _benchmark_shape_parity(16, 1280, 2, [0, 15])

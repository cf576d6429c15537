"""GPU parity of the attention blocks through the C ABI, against the fixtures written from
the reference modules (tests/golden/attention.npz) and against the oracle on the same bf16-rounded input.
Tolerance: bf16 input / output rounding (2^-8 relative each) -> |err| <= 2e-2 * |ref| + 2e-2 against the fp32 fixture,
1e-2 * |ref| + 1e-2 against the oracle fed the bf16-rounded input."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "attention.npz")


def test_cbam_matches_reference_fixture_and_oracle():
    from oracle import detector_oracle as O
    from transparent_object_detection_b200.attention import CBAM
    g = np.load(GOLD)
    for i, (b, c, h, w) in enumerate(g["cbam_cases"]):
        m = CBAM(int(c))
        m.load_state_dict({k: torch.from_numpy(g[f"cbam{i}_{k}"]) for k in ("fc1.weight", "fc2.weight", "conv.weight")})
        x = torch.from_numpy(g[f"cbam{i}_x"])
        y = m(x.cuda()).cpu()
        ref = torch.from_numpy(g[f"cbam{i}_y"])
        assert float(((y - ref).abs() > 2e-2 * ref.abs() + 2e-2).float().mean()) == 0.0, i
        sd = {"m." + k: torch.from_numpy(g[f"cbam{i}_{k}"]) for k in ("fc1.weight", "fc2.weight", "conv.weight")}
        with torch.no_grad():
            ref2 = O.cbam(sd, "m", x.to(torch.bfloat16).float())
        assert float(((y - ref2).abs() > 1e-2 * ref2.abs() + 1e-2).float().mean()) == 0.0, i


def test_cbam_in_place_channel_window_and_large_plane():
    """Channel window of a wider buffer (pitch 96, 64 channels), in place, on a plane that spans several pooling chunks."""
    from oracle import detector_oracle as O
    from transparent_object_detection_b200.attention import cbam_nhwc
    g = torch.Generator().manual_seed(5)
    B, H, W, C = 2, 80, 72, 64
    buf = torch.randn((B, H, W, 96), generator=g).to(torch.bfloat16).cuda()
    keep = buf.clone()
    sd = {"m.fc1.weight": torch.randn((4, C, 1, 1), generator=g) * 0.3, "m.fc2.weight": torch.randn((C, 4, 1, 1), generator=g) * 0.5,
          "m.conv.weight": torch.randn((1, 2, 7, 7), generator=g) * 0.2}
    cbam_nhwc(buf, sd["m.fc1.weight"].flatten(1).cuda(), sd["m.fc2.weight"].flatten(1).cuda(), sd["m.conv.weight"][0].contiguous().cuda(),
              out=buf, channels=C)
    torch.cuda.synchronize()
    assert torch.equal(buf[..., C:], keep[..., C:])                     # the neighbouring channels are untouched
    with torch.no_grad():
        ref = O.cbam(sd, "m", keep[..., :C].float().cpu().permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    err = (buf[..., :C].float().cpu() - ref).abs()
    assert float((err > 1e-2 * ref.abs() + 1e-2).float().mean()) == 0.0


def _sa_bf16_emulation(sd, x, fused=False):
    """The oracle's SelfAttention with the build's rounding points: bf16 input, q / k / gamma-scaled v rounded to bf16,
    f32 scores, bf16 attention weights, f32 accumulation.  fused: the query projection carries log2(e) before it is
    rounded to bf16 (tod_attention_fused evaluates the softmax with base-2 exponentials)."""
    import math
    import torch.nn.functional as F
    bf = lambda t: t.to(torch.bfloat16).float()
    b, c, h, w = x.shape
    xb = bf(x)
    qs = math.log2(math.e) if fused else 1.0
    q = bf(F.conv2d(xb, bf(qs * sd["query.weight"]), qs * sd["query.bias"])).view(b, -1, h * w).permute(0, 2, 1) / qs
    k = bf(F.conv2d(xb, bf(sd["key.weight"]), sd["key.bias"])).view(b, -1, h * w)
    g = float(sd["gamma"])
    v = bf(F.conv2d(xb, bf(g * sd["value.weight"]), None)).view(b, -1, h * w)
    att = bf(torch.softmax(torch.bmm(q, k), dim=-1))
    out = torch.bmm(v, att.permute(0, 2, 1)).view(b, c, h, w)
    return out + (g * sd["value.bias"]).view(1, -1, 1, 1) + xb


@pytest.mark.parametrize("fused", [True, False], ids=["fused", "unfused"])
def test_self_attention_matches_reference_fixture_and_emulation(fused):
    """SelfAttention (unfused tcgen05 GEMMs + row softmax) against the fixture written from the reference module (gamma
    != 0): bf16 q / k move the logits by ~|s| * 2^-8, so the stated tolerance against fp32 is 6e-2 * |ref| + 6e-2; against
    the emulation with the same rounding points it is 2e-2."""
    from transparent_object_detection_b200.attention import SelfAttention
    g = np.load(GOLD)
    for i, (b, c, h, w) in enumerate(g["sa_cases"]):
        keys = ("query.weight", "query.bias", "key.weight", "key.bias", "value.weight", "value.bias", "gamma")
        sd = {k: torch.from_numpy(g[f"sa{i}_{k}"]) for k in keys}
        m = SelfAttention(int(c))
        m.fused = fused
        m.load_state_dict(sd)
        x = torch.from_numpy(g[f"sa{i}_x"])
        y = m(x.cuda()).cpu()
        ref = torch.from_numpy(g[f"sa{i}_y"])
        assert float(((y - ref).abs() > 6e-2 * ref.abs() + 6e-2).float().mean()) <= 2e-3, (i, float((y - ref).abs().max()))
        emu = _sa_bf16_emulation(sd, x, fused)
        assert float(((y - emu).abs() > 2e-2 * emu.abs() + 2e-2).float().mean()) == 0.0, (i, float((y - emu).abs().max()))
        assert float((y - x).abs().max()) > 0.1                      # gamma != 0: the attention term is really there


def test_softmax_rows_against_torch():
    from transparent_object_detection_b200 import lib
    from transparent_object_detection_b200._lib import check
    g = torch.Generator().manual_seed(2)
    for rows, cols, pitch in [(7, 96, 96), (3, 6400, 6400), (5, 48, 64)]:
        s = (torch.randn((rows, pitch), generator=g) * 6).cuda()
        o = torch.zeros((rows, pitch), dtype=torch.bfloat16, device="cuda")
        check(lib().tod_softmax_rows_f32_bf16(s.data_ptr(), o.data_ptr(), rows, cols, pitch, pitch,
                                              torch.cuda.current_stream().cuda_stream), "softmax")
        torch.cuda.synchronize()
        want = torch.softmax(s[:, :cols], -1)
        err = (o[:, :cols].float() - want).abs()
        assert float((err > 1e-2 * want + 1e-6).float().mean()) == 0.0
        assert abs(float(o[:, :cols].float().sum(1).mean()) - 1.0) < 5e-3


@pytest.mark.parametrize("fused", [True, False], ids=["fused", "unfused"])
def test_self_attention_at_network_size(fused):
    """The backbone's instance at scale s, 640x640 (model/backbone.py:33): C = 128 on the 80x80 map, N = 6400 tokens, one
    image; against the emulation with the same rounding points (f32 scores are 164 MB, bf16 weights 82 MB per image)."""
    from transparent_object_detection_b200.attention import SelfAttention
    g = torch.Generator().manual_seed(8)
    c, h, w = 128, 80, 80
    m = SelfAttention(c)
    m.fused = fused
    sd = {"query.weight": torch.randn((16, c, 1, 1), generator=g) * 0.08, "query.bias": torch.randn((16,), generator=g) * 0.1,
          "key.weight": torch.randn((16, c, 1, 1), generator=g) * 0.08, "key.bias": torch.randn((16,), generator=g) * 0.1,
          "value.weight": torch.randn((c, c, 1, 1), generator=g) * 0.1, "value.bias": torch.randn((c,), generator=g) * 0.1,
          "gamma": torch.tensor([0.5])}
    m.load_state_dict(sd)
    x = torch.randn((1, c, h, w), generator=g)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    xc = x.cuda()
    m(xc)                                   # warm-up (function attributes, module loading)
    e0.record()
    y = m(xc)
    e1.record()
    torch.cuda.synchronize()
    emu = _sa_bf16_emulation(sd, x, fused)
    err = (y.cpu() - emu).abs()
    assert float((err > 2e-2 * emu.abs() + 2e-2).float().mean()) == 0.0, float(err.max())
    print(f"SelfAttention 1 x 128 x 80 x 80 (N = 6400), {'fused' if fused else 'unfused'}, layout conversions included: "
          f"{e0.elapsed_time(e1):.2f} ms")


def test_current_source_network_matches_reference_fixture():
    """BaseModel(..., attention=True): the CURRENT-SOURCE backbone and head (2 + 12 CBAM, 1 SelfAttention with gamma = 0.5)
    around the plain neck, scale n, 2 x 3 x 64 x 96, against the reference's own modules (fixture written by
    oracle/make_golden_attention.py): stage features within 4 % of abs-max, boxes <= 1.5 px, scores <= 8e-3; the captured
    graph (Detector) gives the rows of the oracle's NMS on the eager decoded tensor."""
    from oracle import detector_oracle as O, synth
    from transparent_object_detection_b200 import BaseModel, DecodeBox, Detector
    g = np.load(os.path.join(os.path.dirname(GOLD), "net_n_attention_64x96.npz"))
    C_, d, m = synth.SCALES["n"]
    sd = synth.make_state_dict(80, C_, d, m, seed=0)
    sd.update(synth.make_attention_state_dict(80, C_, d, m, seed=0))
    model = BaseModel(80, C_, d, m, attention=True).eval()
    missing = model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    x = torch.from_numpy(synth.make_images(2, 64, 96, seed=7))
    out = model(x.cuda()).cpu()
    eng = model.engine(2, 64, 96, torch.device("cuda", torch.cuda.current_device()))
    assert sum(1 for k, _, _ in eng.ops if k == "cbam") == 14 and sum(1 for k, _, _ in eng.ops if k == "attn") == 1
    for name in ("p3", "p4", "p5"):
        ref = torch.from_numpy(g[name])
        err = (eng.feature_nchw(name).cpu() - ref).abs()
        assert float(err.max()) <= 0.04 * float(ref.abs().max()), (name, float(err.max()), float(ref.abs().max()))
    want = torch.from_numpy(g["out"])
    assert float((out[:, :4] - want[:, :4]).abs().max()) <= 1.5
    assert float((out[:, 4:] - want[:, 4:]).abs().max()) <= 8e-3
    # and it differs from the plain topology (the attention blocks are really in the path)
    plain = BaseModel(80, C_, d, m).eval()
    plain.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.make_state_dict(80, C_, d, m, seed=0).items()})
    assert float((plain(x.cuda()).cpu() - out).abs().max()) > 0.5
    db = DecodeBox(80, (64, 96))
    dec = db.decode_box(model(x.cuda()))
    conf = float(dec[:, :, 4:].max()) * 0.5
    want_rows = O.non_max_suppression(dec.cpu().numpy().copy(), 80, (64, 96), (64, 96), True, conf, 0.5)
    got_rows = Detector(model, (64, 96), confidence=conf, nms_iou=0.5).detect(x)
    assert any(r is not None for r in want_rows)
    for a, b in zip(got_rows, want_rows):
        assert (a is None) == (b is None) and (a is None or np.array_equal(a, b))


@pytest.mark.parametrize("c,h,w", [(192, 8, 12), (256, 12, 12), (64, 16, 16), (384, 4, 8)])
def test_self_attention_other_widths_against_oracle(c, h, w):
    """The instance's width at the other scales: 4C = 64 (n), 192 (m: q/k width 24 -> 32, two K steps, 64-byte swizzle),
    256 (l), 384 (x: above the fused kernel's 256 channels -> the unfused GEMM chain), against the oracle on the
    bf16-rounded input and the emulation with the build's rounding points."""
    from oracle import detector_oracle as O
    from transparent_object_detection_b200.attention import SelfAttention
    g = torch.Generator().manual_seed(c)
    d = c // 8
    sd = {"query.weight": torch.randn((d, c, 1, 1), generator=g) * (0.7 / c ** 0.5), "query.bias": torch.randn((d,), generator=g) * 0.1,
          "key.weight": torch.randn((d, c, 1, 1), generator=g) * (0.7 / c ** 0.5), "key.bias": torch.randn((d,), generator=g) * 0.1,
          "value.weight": torch.randn((c, c, 1, 1), generator=g) * (1.0 / c ** 0.5), "value.bias": torch.randn((c,), generator=g) * 0.1,
          "gamma": torch.tensor([0.8])}
    m = SelfAttention(c)
    m.load_state_dict(sd)
    x = torch.randn((2, c, h, w), generator=g)
    y = m(x.cuda()).cpu()
    fused = c % 32 == 0 and c <= 256
    emu = _sa_bf16_emulation(sd, x, fused)
    assert float(((y - emu).abs() > 2e-2 * emu.abs() + 2e-2).float().mean()) == 0.0, float((y - emu).abs().max())
    with torch.no_grad():
        ref = O.self_attention({"m." + k: v for k, v in sd.items()}, "m", x.to(torch.bfloat16).float())
    assert float(((y - ref).abs() > 6e-2 * ref.abs() + 6e-2).float().mean()) <= 2e-3, float((y - ref).abs().max())


def test_unfused_self_attention_changing_inputs_with_poisoned_temporaries():
    """in the unfused chain the "weight" operand of the scores / value^T / output GEMMs is an activation the
    previous kernel has just written (TOD_CONV_DYNAMIC_W).  A kernel that prefetched it ahead of its programmatic-launch
    wait would read the PREVIOUS call's (or poisoned) memory: two different inputs back to back, the allocator's free
    blocks filled with NaN in between, each result checked against the emulation of ITS input."""
    from transparent_object_detection_b200.attention import SelfAttention
    g = torch.Generator().manual_seed(21)
    c, h, w = 64, 16, 24              # q / k width 8 -> padded to 16, cout = N = 384 > 128: the resident-weight plan
    m = SelfAttention(c)
    m.fused = False
    sd = {"query.weight": torch.randn((c // 8, c, 1, 1), generator=g) * 0.1, "query.bias": torch.randn((c // 8,), generator=g) * 0.1,
          "key.weight": torch.randn((c // 8, c, 1, 1), generator=g) * 0.1, "key.bias": torch.randn((c // 8,), generator=g) * 0.1,
          "value.weight": torch.randn((c, c, 1, 1), generator=g) * 0.1, "value.bias": torch.randn((c,), generator=g) * 0.1,
          "gamma": torch.tensor([0.7])}
    m.load_state_dict(sd)
    for it in range(4):
        x = torch.randn((2, c, h, w), generator=g) * (1.0 + it)
        junk = [torch.full((n,), float("nan"), dtype=torch.float32, device="cuda") for n in (1 << 12, 1 << 16, 1 << 20, 1 << 22)]
        torch.cuda.synchronize()
        del junk                                      # freed blocks keep their NaN contents and are handed out again
        y = m(x.cuda()).cpu()
        emu = _sa_bf16_emulation(sd, x, False)
        assert torch.isfinite(y).all(), it
        assert float(((y - emu).abs() > 2e-2 * emu.abs() + 2e-2).float().mean()) == 0.0, (it, float((y - emu).abs().max()))

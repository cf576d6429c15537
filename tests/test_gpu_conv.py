"""GPU parity of the tcgen05 implicit-GEMM conv (through the C ABI) against an fp32 torch evaluation of the
same op on the same bf16-rounded operands, and against the CUDA-core check kernel.
Tolerance: bf16 output rounding (2^-8 relative) + fp32 accumulation-order noise -> |err| <= 2e-2*|ref| + 2e-2
for bf16 outputs, 2e-3*|ref| + 2e-3 for f32 outputs (operands are O(1), K up to 4608)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

# (name, B, H, W, cin, cout, k, stride, opts)
CASES = [
    ("1x1_k64_n64", 1, 16, 16, 64, 64, 1, 1, {}),
    ("1x1_k256_n128", 2, 20, 20, 256, 128, 1, 1, {}),
    ("1x1_k96_bk32_views", 2, 12, 16, 96, 64, 1, 1, {"x_pitch": 160, "x_off": 32, "out_pitch": 192, "out_off": 64}),
    ("1x1_k48_bk16", 1, 8, 8, 48, 32, 1, 1, {}),
    ("1x1_n512_two_ntiles", 1, 20, 20, 128, 512, 1, 1, {}),
    ("1x1_n384_two_ntiles", 1, 10, 12, 64, 384, 1, 1, {}),
    ("1x1_n80_f32_noact", 2, 20, 20, 128, 80, 1, 1, {"f32": True, "act": 0}),
    ("1x1_small_m", 1, 3, 5, 64, 32, 1, 1, {}),
    ("1x1_upadd", 2, 8, 12, 64, 64, 1, 1, {"upadd": True}),
    ("1x1_nobias_f32", 1, 8, 8, 64, 32, 1, 1, {"f32": True, "act": 0, "bias": False}),
    ("3x3_c64_20x20", 2, 20, 20, 64, 64, 3, 1, {}),
    ("3x3_c32_res", 2, 16, 24, 32, 32, 3, 1, {"residual": True}),
    ("3x3_c128_40x40", 1, 40, 40, 128, 128, 3, 1, {}),
    ("3x3_c16", 1, 12, 16, 16, 16, 3, 1, {}),
    ("3x3_views_res", 1, 20, 20, 64, 64, 3, 1, {"x_pitch": 256, "x_off": 64, "out_pitch": 256, "out_off": 128,
                                                  "residual": True}),
    ("3x3_k4608_n64", 1, 20, 20, 512, 64, 3, 1, {}),
    ("3x3_n256", 1, 20, 20, 64, 256, 3, 1, {}),
    ("3x3s2_c64_n128", 2, 32, 32, 64, 128, 3, 2, {}),
    ("3x3s2_c32_n64_rect", 1, 40, 24, 32, 64, 3, 2, {}),
    ("3x3s2_c256_n512", 1, 40, 40, 256, 512, 3, 2, {}),
    ("3x3_stages1", 1, 16, 16, 64, 64, 3, 1, {"stages": 1}),
    ("3x3_stages2", 1, 16, 16, 64, 64, 3, 1, {"stages": 2}),
    ("3x3_big_160", 1, 160, 160, 32, 32, 3, 1, {}),
    ("3x3_40x40_partial_rows", 2, 40, 40, 64, 64, 3, 1, {"residual": True}),
    ("3x3_tiny_6x6", 3, 6, 6, 32, 48, 3, 1, {}),
    ("3x3_c256_n128_ring", 1, 24, 24, 256, 128, 3, 1, {}),
    ("3x3_c128_n128_m1", 2, 16, 16, 128, 128, 3, 1, {"m": 1}),
    ("3x3_c64_ring", 2, 32, 16, 64, 64, 3, 1, {"no_station": 1}),
    ("3x3s2_c128_n128", 2, 48, 32, 128, 128, 3, 2, {}),
    ("3x3s2_c16_n32_odd_tiles", 1, 20, 28, 16, 32, 3, 2, {}),
    ("1x1_k512_n256_ring", 1, 40, 40, 512, 256, 1, 1, {}),
    ("1x1_k768_n512", 1, 20, 20, 768, 512, 1, 1, {}),
    ("1x1_n144_f32", 1, 16, 16, 64, 144, 1, 1, {"f32": True, "act": 0}),
    ("1x1_m_not_mult_128", 3, 7, 9, 32, 32, 1, 1, {}),
    ("1x1_res_views", 2, 16, 16, 64, 64, 1, 1, {"residual": True, "out_pitch": 128, "out_off": 64}),
    # upsample-add operand through the shared-memory ring (1x1 on 16x8 pixel tiles): ragged tiles, several panels, weights
    # that do not stay resident, two N tiles, channel windows on both sides
    ("1x1_upadd_ragged_40x40_n256", 2, 40, 40, 256, 256, 1, 1, {"upadd": True}),
    ("1x1_upadd_80x80_n128_views", 1, 80, 80, 128, 128, 1, 1, {"upadd": True, "x_pitch": 192, "x_off": 64, "out_pitch": 384,
                                                                 "out_off": 128}),
    ("1x1_upadd_n512_two_ntiles", 1, 20, 12, 64, 512, 1, 1, {"upadd": True}),
    ("1x1_upadd_n32_register_path", 1, 12, 12, 64, 32, 1, 1, {"upadd": True}),
    ("1x1_upadd_noact_register_path", 1, 16, 16, 64, 64, 1, 1, {"upadd": True, "act": 0}),
    # row-flat tiles (3x3 stride 1 on maps that 16x8 patches tile badly): one tile per image, ragged last tile, residual
    # through the ring and through registers, channel windows, N tiles of 128 chosen by the plan, narrow K chunks
    ("3x3_flat_10x10_n96", 3, 10, 10, 64, 96, 3, 1, {}),
    ("3x3_flat_18x20_res_views", 2, 18, 20, 64, 64, 3, 1, {"x_pitch": 192, "x_off": 64, "out_pitch": 160, "out_off": 32,
                                                           "residual": True}),
    ("3x3_flat_20x20_n512", 1, 20, 20, 128, 512, 3, 1, {}),
    ("3x3_40x40_c256_res", 1, 40, 40, 256, 256, 3, 1, {"residual": True}),
    ("3x3_flat_20x12_rect_c32", 2, 20, 12, 32, 32, 3, 1, {"residual": True}),
    ("3x3_flat_17x9_c16_f32", 1, 17, 9, 16, 48, 3, 1, {"f32": True, "act": 0}),
    ("3x3_flat_126wide", 1, 4, 126, 32, 64, 3, 1, {}),
]
VARIANTS = {"pertap": 1, "halo": 2}


def make_case(B, H, W, cin, cout, k, stride, opts, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x_pitch, x_off = opts.get("x_pitch", cin), opts.get("x_off", 0)
    out_pitch, out_off = opts.get("out_pitch", cout), opts.get("out_off", 0)
    Ho, Wo = H // stride, W // stride
    x = torch.randn((B, H, W, x_pitch), generator=g).to(torch.bfloat16).cuda()
    w = torch.randn((cout, cin, k, k), generator=g) * (2.0 / (cin * k * k)) ** 0.5
    bias = torch.randn((cout,), generator=g) * 0.5 if opts.get("bias", True) else None
    dt = torch.float32 if opts.get("f32") else torch.bfloat16
    out = torch.full((B, Ho, Wo, out_pitch), 7.0, dtype=dt).cuda()
    res = torch.randn((B, Ho, Wo, out_pitch), generator=g).to(torch.bfloat16).cuda() if opts.get("residual") else None
    up = torch.randn((B, Ho // 2, Wo // 2, cout), generator=g).cuda() if opts.get("upadd") else None
    return dict(x=x, x_off=x_off, cin=cin, w=w, bias=bias, out=out, out_off=out_off, stride=stride,
                act=opts.get("act", 1), res=res, res_off=out_off if res is not None else 0, up=up,
                stages=opts.get("stages", 0), cout=cout, m=opts.get("m", 0), no_station=opts.get("no_station", 0))


def run_case(case, simt=False, variant=0):
    from tests import gpu_util as U
    out = torch.full_like(case["out"], 7.0)
    U.run_conv(case["x"], case["x_off"], case["cin"], case["w"], case["bias"], out, case["out_off"], case["stride"],
               case["act"], case["res"], case["res_off"], case["up"], 0, case["stages"], simt=simt, variant=variant,
               m=case["m"], no_station=case["no_station"])
    want = U.conv_reference(case["x"], case["x_off"], case["cin"], case["w"], case["bias"], case["stride"], case["act"],
                            case["res"], case["res_off"], case["up"])
    got = out[..., case["out_off"]:case["out_off"] + case["cout"]].float()
    return got, want, out


@pytest.mark.parametrize("variant", list(VARIANTS), ids=list(VARIANTS))
@pytest.mark.parametrize("case_def", CASES, ids=[c[0] for c in CASES])
def test_conv_tcgen05_matches_fp32_reference(case_def, variant):
    from tests import gpu_util as U
    name, B, H, W, cin, cout, k, s, opts = case_def
    case = make_case(B, H, W, cin, cout, k, s, opts)
    got, want, out = run_case(case, variant=VARIANTS[variant])
    f32 = bool(opts.get("f32"))
    rep = U.error_report(got, want, name, 2e-3 if f32 else 2e-2, 2e-3 if f32 else 2e-2)
    assert rep["bad_frac"] == 0 and rep["nan"] == 0, rep
    # channels outside the output window are untouched (concat-by-offset contract)
    off, c = case["out_off"], cout
    if out.shape[-1] > c:
        mask = torch.ones(out.shape[-1], dtype=torch.bool, device=out.device)
        mask[off:off + c] = False
        assert bool((out[..., mask].float() == 7.0).all())


@pytest.mark.parametrize("case_def", CASES[:4] + CASES[10:12] + CASES[17:19], ids=lambda c: c[0])
def test_simt_check_kernel_matches_reference(case_def):
    from tests import gpu_util as U
    name, B, H, W, cin, cout, k, s, opts = case_def
    case = make_case(B, H, W, cin, cout, k, s, opts)
    got, want, _ = run_case(case, simt=True)
    f32 = bool(opts.get("f32"))
    rep = U.error_report(got, want, name, 2e-3 if f32 else 2e-2, 2e-3 if f32 else 2e-2)
    assert rep["bad_frac"] == 0 and rep["nan"] == 0, rep


@pytest.mark.parametrize("case_def", [c for c in CASES if c[6] == 3 and c[7] == 1], ids=lambda c: c[0])
def test_flat_tiles_reverse_order_and_patch_tiles_are_bit_identical(case_def):
    """The tile geometry (row-flat tiles vs 16x8 patches, TOD_CONV_PATCH_TILES) and the tile order (TOD_CONV_REVERSE) only
    change WHICH CTA computes a pixel and when: the MMA sequence per output element is the same, so the outputs of the halo
    kernel must be bit-identical in all four combinations."""
    from tests import gpu_util as U
    from transparent_object_detection_b200._lib import TOD_CONV_PATCH_TILES, TOD_CONV_REVERSE
    name, B, H, W, cin, cout, k, s, opts = case_def
    case = make_case(B, H, W, cin, cout, k, s, opts)
    outs = []
    for flags in (0, TOD_CONV_PATCH_TILES, TOD_CONV_REVERSE, TOD_CONV_PATCH_TILES | TOD_CONV_REVERSE):
        out = torch.full_like(case["out"], 7.0)
        U.run_conv(case["x"], case["x_off"], case["cin"], case["w"], case["bias"], out, case["out_off"], case["stride"],
                   case["act"], case["res"], case["res_off"], case["up"], 0, case["stages"], variant=2, m=case["m"],
                   no_station=case["no_station"], flags=flags)
        outs.append(out)
    for o in outs[1:]:
        assert torch.equal(o, outs[0])


@pytest.mark.parametrize("reverse", [0, 1], ids=["fwd", "rev"])
@pytest.mark.parametrize("case_def", CASES, ids=[c[0] for c in CASES])
def test_cta_pair_kernel_is_bit_identical_to_single_cta(case_def, reverse):
    """tcgen05 cta_group::2 (TOD_CONV_PAIR_ON): two CTAs share every weight tile as one M = 256 MMA.  Each output element is
    the same sequence of products accumulated in the same order as in the single-CTA halo kernel, so the outputs must be
    bit-identical -- for every tile geometry, an odd number of sub-tiles (the odd CTA idles), residual / upsample-add
    operands, stride 2, both tile orders; channels outside the output window stay untouched."""
    from tests import gpu_util as U
    from transparent_object_detection_b200._lib import TOD_CONV_PAIR_OFF, TOD_CONV_PAIR_ON, TOD_CONV_REVERSE
    name, B, H, W, cin, cout, k, s, opts = case_def
    case = make_case(B, H, W, cin, cout, k, s, opts)
    # the K chunk width fixes the accumulation order; where the two plans would pick different widths by themselves (the
    # pair's half-size weight slots fit 64-channel chunks, the single CTA's do not) it is pinned for both
    bk = {"3x3s2_c256_n512": 32}.get(name, 0)
    outs = []
    for flags in (TOD_CONV_PAIR_OFF, TOD_CONV_PAIR_ON):
        out = torch.full_like(case["out"], 7.0)
        U.run_conv(case["x"], case["x_off"], case["cin"], case["w"], case["bias"], out, case["out_off"], case["stride"],
                   case["act"], case["res"], case["res_off"], case["up"], bk, case["stages"], variant=2, m=case["m"],
                   no_station=case["no_station"], flags=flags | (TOD_CONV_REVERSE if reverse else 0))
        outs.append(out)
    assert torch.equal(outs[0], outs[1])
    got = outs[1][..., case["out_off"]:case["out_off"] + cout].float()
    want = U.conv_reference(case["x"], case["x_off"], case["cin"], case["w"], case["bias"], case["stride"], case["act"],
                            case["res"], case["res_off"], case["up"])
    f32 = bool(opts.get("f32"))
    rep = U.error_report(got, want, name, 2e-3 if f32 else 2e-2, 2e-3 if f32 else 2e-2)
    assert rep["bad_frac"] == 0 and rep["nan"] == 0, rep


def test_conv_rejects_bad_arguments():
    import ctypes as C
    from transparent_object_detection_b200 import _lib
    d = _lib.ConvDesc()
    rc = _lib.lib().tod_conv2d_nhwc_bf16(C.byref(d), None)
    assert rc == -1 and b"null" in _lib.lib().tod_last_error()
    x = torch.zeros((1, 8, 8, 24), dtype=torch.bfloat16).cuda()
    o = torch.zeros((1, 8, 8, 32), dtype=torch.bfloat16).cuda()
    w = torch.zeros((32, 24), dtype=torch.bfloat16).cuda()
    d.d_x, d.d_w, d.d_out = x.data_ptr(), w.data_ptr(), o.data_ptr()
    d.batch, d.hin, d.win, d.cin, d.cout, d.ksize, d.stride = 1, 8, 8, 24, 32, 1, 1
    d.x_pitch, d.out_pitch = 24, 32
    assert _lib.lib().tod_conv2d_nhwc_bf16(C.byref(d), None) == -1      # cin % 16 != 0
    d.cin, d.ksize = 16, 5
    assert _lib.lib().tod_conv2d_nhwc_bf16(C.byref(d), None) == -1      # unsupported kernel size


# ------------------------------------------------------------------------------------------ fused 1x1 tail (back-to-back GEMM)
TAIL_CASES = [
    ("3x3_flat_20x20_c64", 2, 20, 20, 64, 3, 1),    # row-flat tiles under the fused tail (head.box.2.2 geometry)
    ("3x3_flat_10x10_c128", 3, 10, 10, 128, 3, 1),
    ("3x3s2_c32_160", 2, 160, 160, 32, 3, 2),      # backbone.dark2.0 -> dark2.1.cv1 geometry (scale s)
    ("3x3s2_c32_ragged", 1, 40, 56, 32, 3, 2),
    ("3x3_c64", 2, 24, 40, 64, 3, 1),
    ("1x1_c128", 1, 20, 20, 128, 1, 1),
    ("3x3_c16_tiny", 3, 6, 6, 16, 3, 1),
]


@pytest.mark.parametrize("case_def", TAIL_CASES, ids=[c[0] for c in TAIL_CASES])
def test_conv_tail1x1_equals_two_separate_convs(case_def):
    """tod_conv2d_tail1x1 (conv -> SiLU -> bf16 panel in shared memory -> 1x1 conv -> SiLU) against the same two convs
    through tod_conv2d_nhwc_bf16 with the intermediate in HBM: same rounding points, so the outputs must be bit-identical;
    channel-window output (pitch 96, offset 32) like the C2f concat buffer."""
    import ctypes as C
    from tests import gpu_util as U
    from transparent_object_detection_b200 import _lib
    from transparent_object_detection_b200._lib import ConvDesc, ConvTailDesc, check
    from transparent_object_detection_b200.engine import pack_conv_weight
    _, B, H, W, cin, k, stride = case_def
    g = torch.Generator(device="cpu").manual_seed(3)
    x = torch.randn((B, H, W, cin), generator=g).to(torch.bfloat16).cuda()
    w1 = torch.randn((64, cin, k, k), generator=g) * (2.0 / (cin * k * k)) ** 0.5
    b1 = torch.randn((64,), generator=g) * 0.5
    w2 = torch.randn((64, 64, 1, 1), generator=g) * (2.0 / 64) ** 0.5
    b2 = torch.randn((64,), generator=g) * 0.5
    Ho, Wo = H // stride, W // stride
    mid = torch.zeros((B, Ho, Wo, 64), dtype=torch.bfloat16, device="cuda")
    want = torch.full((B, Ho, Wo, 96), 7.0, dtype=torch.bfloat16, device="cuda")
    U.run_conv(x, 0, cin, w1, b1, mid, 0, stride, 1, variant=2)
    U.run_conv(mid, 0, 64, w2, b2, want, 32, 1, 1, variant=2)
    got = torch.full_like(want, 7.0)
    wp1, wp2 = pack_conv_weight(w1, 0).cuda(), pack_conv_weight(w2, 0).cuda()
    bb1, bb2 = b1.cuda().contiguous(), b2.cuda().contiguous()
    d = ConvDesc()
    d.d_x, d.d_w, d.d_bias = x.data_ptr(), wp1.data_ptr(), bb1.data_ptr()
    d.batch, d.hin, d.win, d.cin, d.cout, d.ksize, d.stride = B, H, W, cin, 64, k, stride
    d.x_pitch, d.out_pitch, d.act, d.out_dtype = cin, 64, 1, 0
    t = ConvTailDesc()
    t.d_w2, t.d_bias2, t.d_out2 = wp2.data_ptr(), bb2.data_ptr(), got.data_ptr() + 32 * 2
    t.cout2, t.out2_pitch, t.act2 = 64, 96, 1
    check(_lib.lib().tod_conv2d_tail1x1(C.byref(d), C.byref(t), U.stream()), "tail")
    torch.cuda.synchronize()
    assert torch.equal(got[..., :32], torch.full_like(got[..., :32], 7.0)) and torch.equal(got[..., 96 - 0:], got[..., 96:])
    assert torch.equal(got, want), float((got.float() - want.float()).abs().max())
    ref = U.conv_reference(U.conv_reference(x, 0, cin, w1, b1, stride, 1).to(torch.bfloat16), 0, 64, w2, b2, 1, 1)
    err = (got[..., 32:].float() - ref).abs()
    assert float((err > 2e-2 + 2e-2 * ref.abs()).float().mean()) == 0.0

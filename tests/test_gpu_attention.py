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

"""Writes tests/golden/attention.npz: outputs of the REFERENCE's CBAM and SelfAttention modules (model/blocks.py:190-254,
imported unmodified from /root/reference) on seeded inputs and weights -- SURVEY.md section 8 row f1.  Authoring container
only.  gamma is set to a non-zero value (the reference initialises it to 0, which would make the test vacuous: SURVEY F10).

usage: python -m oracle.make_golden_attention
"""
import os

import numpy as np
import torch

from oracle import ref_import

CBAM_CASES = [(2, 64, 12, 16), (1, 128, 9, 7), (1, 256, 5, 5), (2, 80, 6, 10)]       # (B, C, H, W); C // 16 hidden channels
SA_CASES = [(2, 128, 8, 12), (1, 64, 4, 12)]                                       # H * W % 16 == 0 (stride-8 maps)


def main():
    ref_import.import_reference()
    from model.blocks import CBAM, SelfAttention
    out = {"cbam_cases": np.array(CBAM_CASES, np.int32), "sa_cases": np.array(SA_CASES, np.int32)}
    g = torch.Generator().manual_seed(99)
    for i, (b, c, h, w) in enumerate(CBAM_CASES):
        m = CBAM(c).eval()
        for p in m.parameters():
            p.data = torch.randn(p.shape, generator=g) * (0.3 if p.dim() == 4 and p.shape[-1] == 7 else 2.0 / p.shape[1] ** 0.5)
        x = torch.randn((b, c, h, w), generator=g) * 1.5
        with torch.no_grad():
            y = m(x)
        out[f"cbam{i}_x"], out[f"cbam{i}_y"] = x.numpy(), y.numpy()
        for k, v in m.state_dict().items():
            out[f"cbam{i}_{k}"] = v.numpy()
    for i, (b, c, h, w) in enumerate(SA_CASES):
        m = SelfAttention(c).eval()
        for n_, p in m.named_parameters():
            p.data = torch.randn(p.shape, generator=g) * (0.1 if n_.endswith("bias") else 1.5 / c ** 0.5)
        m.gamma.data = torch.tensor([0.7])
        x = torch.randn((b, c, h, w), generator=g)
        with torch.no_grad():
            y = m(x)
        out[f"sa{i}_x"], out[f"sa{i}_y"] = x.numpy(), y.numpy()
        for k, v in m.state_dict().items():
            out[f"sa{i}_{k}"] = v.numpy()
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "attention.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()


def network_fixture():
    """Scale n with the CURRENT-SOURCE backbone and head (CBAM x 14, SelfAttention) around the plain neck, 2 x 3 x 64 x 96:
    stage features and the eval head tensor of the reference itself -> tests/golden/net_n_attention_64x96.npz."""
    from oracle import synth
    C, d, m = synth.SCALES["n"]
    sd = synth.make_state_dict(80, C, d, m, seed=0)
    sd.update(synth.make_attention_state_dict(80, C, d, m, seed=0))
    model = ref_import.build_reference_model(80, C, d, m, sd, attention=True)
    x = torch.from_numpy(synth.make_images(2, 64, 96, seed=7))
    with torch.no_grad():
        p3, p4, p5 = model.backbone(x)
        out = model(x)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "net_n_attention_64x96.npz")
    np.savez_compressed(path, p3=p3.numpy(), p4=p4.numpy(), p5=p5.numpy(), out=out.numpy())
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    network_fixture()

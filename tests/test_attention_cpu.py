"""SURVEY.md section 8 row f1 without a GPU: the oracle's restatement of the reference's CBAM / SelfAttention blocks
(model/blocks.py:190-254) against fixtures written from the reference modules themselves (oracle/make_golden_attention.py)."""
import os

import numpy as np
import torch

from oracle import detector_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "attention.npz")


def _sd(g, tag):
    return {k[len(tag) + 1:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(tag + "_") and k not in (tag + "_x", tag + "_y")}


def test_cbam_oracle_matches_reference_fixture():
    g = np.load(GOLD)
    for i in range(len(g["cbam_cases"])):
        sd = {"m." + k: v for k, v in _sd(g, f"cbam{i}").items()}
        with torch.no_grad():
            y = O.cbam(sd, "m", torch.from_numpy(g[f"cbam{i}_x"]))
        np.testing.assert_allclose(y.numpy(), g[f"cbam{i}_y"], rtol=1e-5, atol=1e-6)


def test_self_attention_oracle_matches_reference_fixture():
    g = np.load(GOLD)
    for i in range(len(g["sa_cases"])):
        sd = {"m." + k: v for k, v in _sd(g, f"sa{i}").items()}
        assert float(sd["m.gamma"]) != 0.0
        with torch.no_grad():
            y = O.self_attention(sd, "m", torch.from_numpy(g[f"sa{i}_x"]))
        np.testing.assert_allclose(y.numpy(), g[f"sa{i}_y"], rtol=1e-5, atol=1e-5)


def test_current_source_network_oracle_matches_reference_fixture():
    from oracle import synth
    g = np.load(os.path.join(os.path.dirname(GOLD), "net_n_attention_64x96.npz"))
    C_, d, m = synth.SCALES["n"]
    sd = synth.make_state_dict(80, C_, d, m, seed=0)
    sd.update(synth.make_attention_state_dict(80, C_, d, m, seed=0))
    x = torch.from_numpy(synth.make_images(2, 64, 96, seed=7))
    with torch.no_grad():
        feats = O.backbone(sd, x, d, attention=True)
        out = O.forward(sd, x, 80, d, attention=True)
    for name, t in zip(("p3", "p4", "p5"), feats):
        np.testing.assert_allclose(t.numpy(), g[name], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=1e-4, atol=1e-4)


def test_attention_parameter_table_matches_the_state_dict_layout():
    from oracle import synth
    from transparent_object_detection_b200.model import attention_parameter_table
    for scale in ("n", "s", "m"):
        C_, d, m = synth.SCALES[scale]
        want = synth.attention_shapes(80, C_, d, m)
        got = {k: tuple(s) for k, s, _ in attention_parameter_table(80, C_, d, m)}
        assert got == {k: tuple(v) for k, v in want.items()}

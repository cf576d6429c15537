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

"""CPU: the C-ABI library loads and exports every symbol include/tod.h declares; host-side packing and the
drop-in parameter tree match the reference layout.  No compute calls (there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import transparent_object_detection_b200 as T
from transparent_object_detection_b200 import _lib
from oracle import detector_oracle as O
from oracle import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "tod.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tod_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = T.lib()
    declared = header_symbols()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(_lib.SYMBOLS) == declared
    assert L.tod_version() == 100


def test_library_has_no_libcuda_link_dependency():
    # the .so must load on a box without a driver (this one): cudart is static, the driver entry point for
    # cuTensorMapEncodeTiled is resolved at run time
    import subprocess
    out = subprocess.run(["ldd", T.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out and "libcudart" not in out


def test_argument_validation_without_gpu():
    L = T.lib()
    assert L.tod_conv2d_nhwc_bf16(None, None) == -1
    assert b"null descriptor" in L.tod_last_error()
    assert L.tod_nms_workspace_bytes(0, 10) == 0
    assert L.tod_nms_workspace_bytes(64, 8400) > 64 * 8400 * 13
    bk, cp, kt = _lib.weight_layout(96, 3)
    assert (bk, cp, kt) == (32, 96, 864)
    assert _lib.weight_layout(48, 1) == (16, 48, 48)
    assert _lib.weight_layout(512, 3) == (64, 512, 4608)
    assert _lib.weight_layout(48, 1, 64) == (64, 64, 64)


def test_parameter_tree_matches_reference_key_layout():
    for scale in ("n", "s", "m"):
        C, d, m = synth.SCALES[scale]
        model = T.BaseModel(80, C, d, m)
        want = synth.state_dict_shapes(80, C, d, m)          # pinned to the reference by oracle/make_golden.py
        got = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        assert list(got.keys()) == list(want.keys())
        assert got == dict(want)
    sd = synth.make_state_dict(80, 16, 1, 1.0)
    res = T.BaseModel(80, 16, 1, 1.0).load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    assert not res.missing_keys and not res.unexpected_keys


def test_fold_and_pack_weights():
    sd = synth.make_state_dict(80, 16, 1, 1.0)
    for prefix in ("backbone.dark2.0", "neck.h3", "head.cls.1.2"):
        w, b = T.fold_conv_bn(sd, prefix)
        wo, bo = O.fold_bn(sd, prefix)                         # reference fuse_conv algebra (pinned by fixtures)
        np.testing.assert_allclose(w.numpy(), wo.numpy(), rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(b.numpy(), bo.numpy(), rtol=1e-5, atol=1e-7)
    w = torch.arange(2 * 24 * 9, dtype=torch.float32).reshape(2, 24, 3, 3) / 64
    p = T.pack_conv_weight(w)                                  # cin 24 -> block_k 16, cin_pad 32
    assert p.shape == (2, 9 * 32) and p.dtype == torch.bfloat16
    p = p.float().reshape(2, 9, 32)
    assert torch.equal(p[:, :, :24], w.permute(0, 2, 3, 1).reshape(2, 9, 24).to(torch.bfloat16).float())
    assert float(p[:, :, 24:].abs().max()) == 0.0


def test_no_cpu_fallback():
    model = T.BaseModel(80, 16, 1, 1.0).eval()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU path"):
        model(torch.zeros(1, 3, 64, 64))


def test_correct_boxes_matches_reference_fixture(golden):
    g = golden("blocks.npz")
    for lb, key in ((True, "cb_letterbox"), (False, "cb_plain")):
        got = T.DecodeBox.correct_boxes(g["cb_xy"].copy(), g["cb_wh"].copy(), (640, 640), (375, 500), lb)
        np.testing.assert_array_equal(got, g[key])


def test_top_boxes_follows_the_reference_selection():
    """Detector.top_boxes == utils/callbacks.py:163-166 (argsort ascending, reversed, first max_boxes)."""
    import numpy as np
    from transparent_object_detection_b200.model import Detector
    det = Detector.__new__(Detector)
    det.max_boxes = 3
    rows = np.array([[0, 0, 1, 1, 0.2, 1], [0, 0, 1, 1, 0.9, 2], [0, 0, 1, 1, 0.5, 0], [0, 0, 1, 1, 0.7, 4], [0, 0, 1, 1, 0.1, 3]], np.float32)
    got = det.top_boxes(rows)
    assert got[:, 4].tolist() == [np.float32(0.9), np.float32(0.7), np.float32(0.5)]
    assert det.top_boxes(None) is None

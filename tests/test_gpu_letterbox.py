"""GPU parity of the device letterbox (csrc/letterbox.cu through tod_letterbox_bicubic_u8) -- SURVEY.md section 8 row f2.
Bit-exact against the fixtures written from the reference's resize_image (Pillow BICUBIC) and against the numpy
restatement at full size; the Detector's raw-image entry points equal the host-letterboxed path row for row."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "letterbox.npz")


def run(img: np.ndarray, w: int, h: int, lb: bool) -> np.ndarray:
    from transparent_object_detection_b200.preprocess import Letterbox
    src = torch.from_numpy(img if img.ndim == 4 else img[None])
    out = torch.zeros((src.shape[0], h, w, 3), dtype=torch.uint8, device="cuda")
    Letterbox((h, w), lb)(src, out)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def test_letterbox_matches_reference_fixtures_bit_exact():
    gold = np.load(GOLD)
    for i, (iw, ih, w, h) in enumerate(gold["cases"]):
        for lb in (0, 1):
            got = run(gold[f"src{i}"], int(w), int(h), bool(lb))[0]
            assert np.array_equal(got, gold[f"dst{i}_lb{lb}"]), (i, lb)


@pytest.mark.parametrize("iw,ih,w,h,n", [(640, 480, 640, 640, 3), (1280, 720, 640, 640, 1), (375, 500, 640, 640, 2),
                                         (640, 640, 640, 640, 2), (1920, 1080, 1280, 1280, 1), (100, 37, 640, 640, 1)])
def test_letterbox_matches_restatement_at_full_size(iw, ih, w, h, n):
    from oracle import letterbox_oracle as LO
    rng = np.random.default_rng(iw * 7 + ih)
    imgs = rng.integers(0, 256, (n, ih, iw, 3), dtype=np.uint8)
    for lb in (True, False):
        got = run(imgs, w, h, lb)
        for j in range(n):
            assert np.array_equal(got[j], LO.resize_image_u8(imgs[j], (w, h), lb)), (j, lb)


def test_detector_raw_images_equal_host_letterboxed_batch():
    import transparent_object_detection_b200 as T
    from oracle import letterbox_oracle as LO, synth
    C_, d, m = synth.SCALES["n"]
    model = T.BaseModel(80, C_, d, m).eval()
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.make_state_dict(80, C_, d, m, seed=0).items()})
    H, W = 96, 128
    det = T.Detector(model, (H, W), confidence=0.01, nms_iou=0.5)
    rng = np.random.default_rng(5)
    sizes = [(150, 100), (150, 100), (64, 200), (128, 96)]                     # (w, h): two equal ones share a launch
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for (w, h) in sizes]
    got = det.detect_images(imgs)
    host = np.stack([LO.resize_image_u8(im, (W, H), True) for im in imgs])
    want = det.detect(torch.from_numpy(host), np.array([[im.shape[0], im.shape[1]] for im in imgs]))
    assert sum(r is not None for r in want) > 0
    for g, w_ in zip(got, want):
        assert (g is None) == (w_ is None)
        if g is not None:
            assert np.array_equal(g, w_)
    one = det.detect_image_rows(imgs[2])
    assert (one is None) == (want[2] is None) and (one is None or np.array_equal(one, want[2]))


def test_letterbox_rejects_bad_arguments():
    from transparent_object_detection_b200.preprocess import Letterbox
    lbx = Letterbox((64, 64))
    with pytest.raises(ValueError):
        lbx(torch.zeros((1, 8, 8, 3), dtype=torch.float32), torch.zeros((1, 64, 64, 3), dtype=torch.uint8, device="cuda"))
    with pytest.raises(ValueError):
        lbx(torch.zeros((1, 8, 8, 3), dtype=torch.uint8), torch.zeros((1, 32, 64, 3), dtype=torch.uint8, device="cuda"))

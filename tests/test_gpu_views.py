"""GPU parity: the serving view's image plumbing (views/object_detection.py:46-85) vs the oracle and the reference goldens."""
import numpy as np
import pytest

from test_gpu_core import _t

pytestmark = pytest.mark.gpu
F = np.float32


@pytest.mark.parametrize("old_wh,padding", [((640, 480), (52, 52, 0, 0)), ((333, 777), (0, 0, 119, 119)), ((416, 416), (0, 0, 0, 0)),
                                            ((1920, 1080), (91, 91, 0, 0))])
def test_restore_predictions_matches_oracle(lib, cuda, old_wh, padding):
    from oracle import views as ov
    from tfmv_b200.views.object_detection import restore_predictions
    rng = np.random.default_rng(20261018 + old_wh[0])
    n = 700
    c = rng.uniform(-0.1, 1.1, (n, 2))
    wh = np.exp(rng.uniform(np.log(1e-3), np.log(0.8), (n, 2)))
    boxes = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(F)
    boxes[::50] = boxes[1::50]                       # duplicates
    boxes[5, 2] = boxes[5, 0] + F(2.0 / old_wh[0])   # a width right at the 2 px limit
    cid = rng.integers(0, 80, n).astype(np.int32)
    sc = rng.random(n).astype(F)
    cl = rng.random((n, 80)).astype(F)
    cf = rng.random((n, 1)).astype(F)
    want = ov.restore_predictions(boxes, cid, sc, cl, cf, (416, 416), padding, old_wh)
    got = restore_predictions(_t(boxes, cuda), _t(cid, cuda), _t(sc, cuda), _t(cl, cuda), _t(cf, cuda), np.int32([416, 416]), padding,
                              np.int32(old_wh))
    assert 0 < want[0].shape[0] < n
    assert str(got[0].dtype) == "torch.int32"
    for g, w in zip(got, want):
        assert np.array_equal(g.cpu().numpy(), w)
    # an empty input stays empty
    e = restore_predictions(np.zeros((0, 4), F), np.zeros((0,), np.int32), np.zeros((0,), F), np.zeros((0, 80), F), np.zeros((0, 1), F),
                            (416, 416), padding, old_wh)
    assert e[0].shape[0] == 0 and e[3].shape[0] == 0


# ---- letterbox (views/object_detection.py:50-62, utils/image_helper.py:293-325) ----
import hashlib  # noqa: E402
import os  # noqa: E402
import sys  # noqa: E402

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from letterbox_inputs import CASES, make_image  # noqa: E402

LB_GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_letterbox.npz"))
_sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("i", range(len(CASES)))
def test_letterbox_reproduces_reference_goldens(lib, cuda, i):
    """The CUDA path against outputs of the reference's own opencvProportionalResize + predict conversion (bit-exact)."""
    from tfmv_b200.ai_models.utils import image_helper as ih
    from tfmv_b200.views.object_detection import prepare_image
    name, h, w, size, kind = CASES[i]
    img_old = make_image(i, h, w, kind)
    bg = tuple(int(v) for v in LB_GOLD[name + "/bg"])
    u8, f32, padding, _ = ih.letterbox(img_old, size, bg, True, True)
    assert tuple(padding) == tuple(int(v) for v in LB_GOLD[name + "/padding"])
    if name + "/u8" in LB_GOLD.files:
        assert np.array_equal(u8.cpu().numpy(), LB_GOLD[name + "/u8"])
    assert _sha(u8.cpu().numpy()) == str(LB_GOLD[name + "/sha_u8"])
    assert _sha(f32.cpu().numpy()[None]) == str(LB_GOLD[name + "/sha_f32"])
    res, pts, pad2 = ih.opencvProportionalResize(img_old, np.int32(size), bg_color=bg)
    assert _sha(res.cpu().numpy()) == str(LB_GOLD[name + "/sha_u8"]) and tuple(pad2) == tuple(padding) and pts.shape == (0,)
    if bg == (0, 0, 0):
        p_img, p_pad, old = prepare_image(img_old, size)
        assert tuple(p_img.shape) == (1, size[1], size[0], 3) and str(p_img.dtype) == "torch.float32"
        assert _sha(p_img.cpu().numpy()) == str(LB_GOLD[name + "/sha_f32"]) and tuple(p_pad) == tuple(padding)
        assert old.tolist() == [w, h]


def test_letterbox_random_sizes_match_oracle(lib, cuda):
    import torch
    from oracle import letterbox as olb
    from tfmv_b200.ai_models.utils import image_helper as ih
    rng = np.random.default_rng(20261018)
    done = 0
    for it in range(120):
        tw, th = int(rng.integers(8, 200)), int(rng.integers(8, 200))
        h, w = int(rng.integers(th, 6 * th)), int(rng.integers(tw, 6 * tw))
        if it % 6 == 1:     # smaller than the target in one or both directions: OpenCV's bilinear fallback
            h, w = int(rng.integers(1, 2 * th)), int(rng.integers(1, tw + 1))
        if it % 5 == 0:     # integral shrink of both sides (block / 2x2 / copy paths)
            k = int(rng.integers(1, 5)); h, w = th * k, tw * k
        if it % 7 == 0:
            h, w = th + int(rng.integers(0, 2)), tw * 2   # one axis at scale 1 after the proportional fit
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8) if it % 3 else (rng.integers(0, 2, (h, w, 3)) * 255).astype(np.uint8)
        bg = tuple(int(v) for v in rng.integers(0, 256, 3))
        try:
            want, pad = olb.proportional_resize(img, (tw, th), bg_color=bg)
        except olb.UnsupportedResize:   # the proportional size truncated to 0: cv2.resize raises as well
            with pytest.raises(ValueError, match="empty"):
                ih.letterbox(img, (tw, th), bg, True, False)
            continue
        src = img if it % 2 else torch.from_numpy(img).pin_memory()   # pageable numpy / pinned host read in place
        if it % 4 == 0:
            src = torch.from_numpy(img).to(cuda)
        u8, f32, padding, (rw, rh) = ih.letterbox(src, (tw, th), bg, True, True)
        assert tuple(padding) == tuple(pad) and (rw, rh) == olb.proportional_size(w, h, tw, th)
        assert np.array_equal(u8.cpu().numpy(), want), (h, w, tw, th)
        assert np.array_equal(f32.cpu().numpy(), want[..., ::-1].astype(np.float32) / 255)
        done += 1
    assert done > 80


def test_letterbox_points_and_refusals(lib, cuda):
    from tfmv_b200.ai_models.utils import image_helper as ih
    img = make_image(0, 480, 640, "noise")
    _, pts, _ = ih.opencvProportionalResize(img, np.int32((96, 96)), points=LB_GOLD["points_in"].tolist(), bg_color=(0, 0, 0))
    assert pts.dtype == np.float32 and np.array_equal(pts, LB_GOLD["points_out"])
    up, _, pad = ih.opencvProportionalResize(np.full((100, 100, 3), 7, np.uint8), (416, 416))   # enlarging works too
    assert tuple(up.shape) == (416, 416, 3) and pad == (0, 0, 0, 0) and int(up.min()) == 7 and int(up.max()) == 7
    with pytest.raises(ValueError, match="empty"):
        ih.opencvProportionalResize(np.zeros((2, 4000, 3), np.uint8), (416, 416))
    with pytest.raises(RuntimeError, match="3-channel"):
        ih.letterbox(np.zeros((500, 500, 4), np.uint8), (416, 416), (0, 0, 0), True, False)
    with pytest.raises(NotImplementedError):
        ih.opencvProportionalResize(img, (96, 96), bg_color=None)
    with pytest.raises(ValueError):
        ih.letterbox(np.zeros((500, 500, 3), np.float32), (416, 416), (0, 0, 0), True, False)

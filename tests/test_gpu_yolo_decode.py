"""GPU parity: GetBoxes / GetNMSBoxes against the oracle (indices bit-exact, values bit-exact via shared detmath)."""
import numpy as np
import pytest

from test_gpu_core import _t, assert_bits_equal

pytestmark = pytest.mark.gpu
F = np.float32


def _check_image(r, b, want, C):
    k = int(r["count"][b])
    assert k == want["selected"].shape[0]
    assert r["sel_idx"][b, :k].tolist() == want["selected"].tolist()
    assert r["sel_anchor"][b, :k].tolist() == want["cand_anchor"][want["selected"]].tolist()
    assert r["classes_id"][b, :k].tolist() == want["classes_id"].tolist()
    assert_bits_equal(r["boxes"][b, :k], want["boxes"])
    assert_bits_equal(r["scores"][b, :k], want["scores"])
    assert_bits_equal(r["classes"][b, :k], want["classes"].reshape(-1, C))
    assert_bits_equal(r["confidence"][b, :k], want["confidence"].reshape(-1, 1))


@pytest.mark.parametrize("image,batch,iou_type,thr", [
    (416, 1, "iou", (0.5, 0.3, 0.5)),      # BASELINE config 1
    (416, 3, "diou", (0.5, 0.2, 0.5)),     # Predict defaults, yolo_v4/model.py:398
    (608, 2, "ciou", (0.5, 0.3, 0.45)),
    (96, 5, "diou", (0.3, 0.3, 0.5)),      # tiny grids: tiles span several images, ragged tail tiles
])
def test_get_nms_boxes_random_init(lib, cuda, image, batch, iou_type, thr):
    from oracle import yolo as oy
    from tfmv_b200 import synth
    from tfmv_b200.ai_models.utils.tf_yolo_utils import GetNMSBoxesBatch
    rng = np.random.default_rng(20261018 + 1 + image + batch)
    heads = synth.yolo_heads(rng, batch, image)
    anc = synth.yolo_anchors()
    r = GetNMSBoxesBatch(*[_t(h, cuda) for h in heads], anc, (image, image), 80, thr[0], thr[1], thr[2], iou_type,
                         with_indices=True)
    r = {k: v.cpu().numpy() for k, v in r.items()}
    for b in range(batch):
        want = oy.get_nms_boxes_ex(*[h[b:b + 1] for h in heads], anc, (image, image), 80, thr[0], thr[1], thr[2], iou_type)
        _check_image(r, b, want, 80)


@pytest.mark.parametrize("iou_type", ["iou", "ciou"])
def test_get_nms_boxes_more_images_than_sms(lib, cuda, iou_type):
    """Batches that need more than one NMS CTA per SM run 512-thread CTAs (two per SM): same results, image by image.
    Heads mix random-init records with planted duplicates so that suppression really happens."""
    import torch
    from oracle import yolo as oy
    from tfmv_b200 import synth
    from tfmv_b200.ai_models.utils.tf_yolo_utils import GetNMSBoxesBatch
    batch = torch.cuda.get_device_properties(0).multi_processor_count + 12
    image = 96
    rng = np.random.default_rng(20261018 + 11)
    heads = synth.yolo_heads_trained_like(rng, batch, image, objects=12, dups=4)
    for h in heads:
        h[::2] = rng.standard_normal(h[::2].shape, dtype=F)  # every other image: pure random-init
    anc = synth.yolo_anchors()
    r = GetNMSBoxesBatch(*[_t(h, cuda) for h in heads], anc, (image, image), 80, 0.5, 0.3, 0.5, iou_type, with_indices=True)
    r = {k: v.cpu().numpy() for k, v in r.items()}
    supp = 0
    for b in range(batch):
        want = oy.get_nms_boxes_ex(*[h[b:b + 1] for h in heads], anc, (image, image), 80, 0.5, 0.3, 0.5, iou_type)
        supp += want["cand_boxes"].shape[0] - want["selected"].shape[0]
        _check_image(r, b, want, 80)
    assert supp > 50


def test_get_nms_boxes_trained_like_and_dropin_signature(lib, cuda):
    from oracle import yolo as oy
    from tfmv_b200 import synth
    from tfmv_b200.ai_models.utils.tf_yolo_utils import GetNMSBoxes, GetNMSBoxesBatch
    rng = np.random.default_rng(7)
    heads = synth.yolo_heads_trained_like(rng, 2, 416)
    anc = synth.yolo_anchors()
    r = GetNMSBoxesBatch(*[_t(h, cuda) for h in heads], anc, (416, 416), 80, 0.5, 0.3, 0.5, "diou", with_indices=True)
    r = {k: v.cpu().numpy() for k, v in r.items()}
    supp = 0
    for b in range(2):
        want = oy.get_nms_boxes_ex(*[h[b:b + 1] for h in heads], anc, (416, 416), 80, 0.5, 0.3, 0.5, "diou")
        supp += want["cand_boxes"].shape[0] - want["selected"].shape[0]
        _check_image(r, b, want, 80)
    assert supp > 10  # duplicates really were suppressed
    # reference call form, batch 1, 5-D heads accepted as well
    h1 = [h[0:1] for h in heads]
    out = GetNMSBoxes(_t(h1[0], cuda), _t(h1[1].reshape(1, 26, 26, 3, 85), cuda), _t(h1[2], cuda),
                      anchors_wh=anc, image_wh=(416, 416), classes_num=80,
                      confidence_thresh=0.5, scores_thresh=0.3, iou_thresh=0.5, iou_type='diou')
    want = oy.get_nms_boxes(*h1, anc, (416, 416), 80, 0.5, 0.3, 0.5, "diou")
    assert len(out) == 5
    assert out[0].shape == want[0].shape and out[3].shape == want[3].shape and out[4].shape == want[4].shape
    assert str(out[1].dtype) == "torch.int32"
    for g, w in zip(out, want):
        if w.dtype == np.int32:
            assert g.cpu().numpy().tolist() == w.tolist()
        else:
            assert_bits_equal(g.cpu().numpy(), w)


def test_get_nms_boxes_edge_cases(lib, cuda):
    from oracle import yolo as oy
    from tfmv_b200 import synth
    from tfmv_b200.ai_models.utils.tf_yolo_utils import GetNMSBoxesBatch
    anc = synth.yolo_anchors()
    rng = np.random.default_rng(9)
    heads = synth.yolo_heads(rng, 2, 96)
    # image 0: nothing passes; image 1: exp overflow -> inf -> 0 -> invalid box dropped (tyu:156,163), saturated
    # sigmoid plateaus (several classes at 1.0 -> first index wins), exact score ties across anchors
    for h in heads:
        h[0, ..., 4::85] = -20.0
    h2 = heads[2].reshape(2, 12, 12, 3, 85)
    h2[1, 0, 0, 0, 2] = 200.0
    h2[1, 0, 0, 0, 4] = 5.0
    h2[1, 0, 1, 1, 5 + 7] = 30.0
    h2[1, 0, 1, 1, 5 + 3] = 25.0
    h2[1, 0, 1, 1, 4] = 5.0
    h2[1, 3, 3, :, 5:] = -3.0
    h2[1, 3, 3, :, 5 + 11] = 1.5
    h2[1, 3, 3, :, 4] = 4.0
    r = GetNMSBoxesBatch(*[_t(h, cuda) for h in heads], anc, (96, 96), 80, 0.5, 0.3, 0.5, "iou", with_indices=True)
    r = {k: v.cpu().numpy() for k, v in r.items()}
    assert r["count"][0] == 0
    want = oy.get_nms_boxes_ex(*[h[1:2] for h in heads], anc, (96, 96), 80, 0.5, 0.3, 0.5, "iou")
    _check_image(r, 1, want, 80)
    assert 3 in want["classes_id"].tolist()  # the plateau case resolved to the first maximal index


@pytest.mark.parametrize("grid", [13, 5])
def test_get_boxes_matches_oracle(lib, cuda, grid):
    from oracle import yolo as oy
    from tfmv_b200.ai_models.utils.tf_yolo_utils import GetBoxes
    rng = np.random.default_rng(31 + grid)
    y = rng.standard_normal((2, grid, grid, 3, 85), dtype=F) * F(2)
    y[0, 1, 1, 0, 2] = 150.0  # inf width -> dropped
    anc = (np.array([[116, 90], [156, 198], [373, 326]], F) / F(416)).astype(F)
    want = oy.get_boxes(y, anc, 80)
    got = GetBoxes(_t(y, cuda), _t(anc, cuda), 80)
    assert got[0].shape[0] == want[0].shape[0] < 2 * grid * grid * 3
    for g, w in zip(got, want):
        assert_bits_equal(g.cpu().numpy(), w)

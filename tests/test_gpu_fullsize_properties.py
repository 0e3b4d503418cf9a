"""Size-independent properties at BASELINE.json's full sizes (the oracle is too slow there):
round trips, linearity over image shards, idempotence of NMS, sortedness, run-to-run determinism."""
import numpy as np
import pytest

from test_gpu_core import _t

pytestmark = pytest.mark.gpu
F = np.float32


def test_config2_yolov4_608_b64_targets_and_loss_properties(lib, cuda):
    import torch
    from tfmv_b200 import synth
    from tfmv_b200.ai_models.datasets.coco_dataset import DataGenerator
    from tfmv_b200.ai_models.utils.tf_yolo_utils import _loss_call
    image, batch = 608, 64
    rng = np.random.default_rng(20261018 + 2)
    anc = synth.yolo_anchors().astype(F)
    heads = [_t(h, cuda) for h in synth.yolo_heads(rng, batch, image)]
    boxes, classes, off = synth.gt_batch(rng, batch, (image, image), max_boxes=100)
    gen = DataGenerator(80, anc, (image, image))
    y_true = gen.GetTargetsBatch(_t(classes, cuda), _t(boxes, cuda), _t(off, cuda))
    # targets: obj in {0,1}; every obj=1 record carries exactly one class bit and its centre lies in its own cell
    n_obj = 0
    for l, t in enumerate(y_true):
        obj = t[..., 4]
        assert bool(((obj == 0) | (obj == 1)).all())
        m = obj == 1
        n_obj += int(m.sum())
        assert bool((t[..., 5:][m].sum(-1) == 1).all())
        assert bool((t[~m] == 0).all())            # nothing but zeros outside object records (collisions cleared)
        b, yy, xx, aa = torch.nonzero(m, as_tuple=True)
        g = t.shape[1]
        assert bool((torch.floor(t[b, yy, xx, aa, 0] * g) == xx).all()) and bool((torch.floor(t[b, yy, xx, aa, 1] * g) == yy).all())
    assert 0.9 * boxes.shape[0] < n_obj <= boxes.shape[0]
    # loss: deterministic, finite, and linear over image shards (the property the multi-GPU path relies on)
    full, parts = _loss_call(y_true, heads, (image, image), anc, 0.5, "ciou", 0, return_parts=True)
    again = _loss_call(y_true, heads, (image, image), anc, 0.5, "ciou", 0)
    assert float(full) == float(again) and np.isfinite(float(full))
    acc = torch.zeros_like(parts)
    for lo, hi in ((0, 24), (24, 64)):
        _, p = _loss_call([t[lo:hi] for t in y_true], [h[lo:hi] for h in heads], (image, image), anc, 0.5, "ciou", 0,
                          batch_divisor=batch, return_parts=True)
        acc += p
    np.testing.assert_allclose(acc.cpu().numpy(), parts.cpu().numpy(), rtol=2e-6)


def test_config1_decode_nms_416_b256_idempotence_and_order(lib, cuda):
    import torch
    from tfmv_b200 import synth
    from tfmv_b200.ai_models.utils.tf_iou_utils import GetIOUNMSByClasses
    from tfmv_b200.ai_models.utils.tf_yolo_utils import GetNMSBoxesBatch
    image, batch = 416, 256
    rng = np.random.default_rng(20261018 + 1)
    heads = [_t(h, cuda) for h in synth.yolo_heads(rng, batch, image)]
    anc = synth.yolo_anchors()
    r = GetNMSBoxesBatch(*heads, anc, (image, image), 80, 0.5, 0.3, 0.5, "iou", with_indices=True)
    r2 = GetNMSBoxesBatch(*heads, anc, (image, image), 80, 0.5, 0.3, 0.5, "iou", with_indices=True)
    cnt = r["count"].cpu().numpy()
    assert (cnt == 500).all()                       # ~5.3k random candidates per image: the cap is reached
    assert torch.equal(r["sel_anchor"], r2["sel_anchor"]) and torch.equal(r["boxes"], r2["boxes"])   # deterministic
    s = r["scores"]
    assert bool((s[:, 1:] <= s[:, :-1]).all())      # emitted in descending score order
    assert bool((r["scores"] > 0.3).all()) and bool((r["confidence"] > 0.5).all())
    b = r["boxes"]
    assert bool((b[..., 2] > b[..., 0]).all()) and bool((b[..., 3] > b[..., 1]).all())
    for k in (0, 100, 255):                          # NMS of its own output keeps everything, in order
        again = GetIOUNMSByClasses(r["boxes"][k], r["scores"][k], r["classes_id"][k], 500, 0.5, "iou")
        assert again.cpu().tolist() == list(range(500))


def test_config3_effdet_d0_round_trip_and_loss_linearity(lib, cuda):
    import torch
    from tfmv_b200 import synth
    from tfmv_b200.ai_models.efficientnet.efficientdet_net_train import get_loss
    from tfmv_b200.ai_models.efficientnet.utils.anchors import Anchors
    c = synth.EFFDET_CONFIGS["d0"]
    a = Anchors(c["min_level"], c["max_level"], c["image_size"], c["num_scales"], c["aspect_ratios"], c["anchor_scale"])
    batch, C = 16, 81
    rng = np.random.default_rng(20261018 + 3)
    boxes, classes, off = synth.gt_batch(rng, batch, (512, 512), max_boxes=100, order="yxyx")
    classes = (classes + 1).astype(np.int32)
    tb, tc, tm = a.generate_targets_batch(_t(boxes, cuda), _t(classes, cuda), _t(off, cuda), C)
    assert sum(int(t.numel()) // 4 for t in tb) == batch * 49104
    # encode -> decode round trip: matched anchors decode back onto one of their image's GT boxes
    dec = a.convert_outputs_boxes(tb)
    gt = _t(boxes, cuda)
    npos = 0
    for l in range(5):
        m = tm[l][..., 0]
        npos += int(m.sum())
        assert bool((tc[l].sum(-1) == 1).all())                      # one-hot everywhere (bg = class 0)
        assert bool((tc[l][..., 0][~m] == 1).all()) and bool((tb[l][~m] == 0).all())
        bi = torch.nonzero(m, as_tuple=True)[0]
        d = dec[l][m]
        for b in range(batch):
            db = d[bi == b]
            if db.numel():
                g = gt[off[b]:off[b + 1]]
                err = (db[:, None, :] - g[None]).abs().amax(-1).amin(-1)
                assert float(err.max()) < 2e-3
    assert npos > 0
    # loss: partial sums are linear over image shards
    pb = [torch.randn_like(t) * 0.25 for t in tb]
    pc = [torch.randn_like(t) for t in tc]
    full, parts, n1 = get_loss(tb, tc, tm, pb, pc, return_parts=True)
    assert float(n1) == npos + 1 and np.isfinite(float(full))
    from tfmv_b200.ai_models.losses.focal_loss import _partial_sums
    s_full, _ = _partial_sums(list(tb), list(tc), list(tm), pb, pc, 0.25, 1.5, 0.1, 0.0)
    s_a, _ = _partial_sums([t[:5] for t in tb], [t[:5] for t in tc], [t[:5] for t in tm], [t[:5] for t in pb], [t[:5] for t in pc], 0.25, 1.5, 0.1, 0.0)
    s_b, _ = _partial_sums([t[5:] for t in tb], [t[5:] for t in tc], [t[5:] for t in tm], [t[5:] for t in pb], [t[5:] for t in pc], 0.25, 1.5, 0.1, 0.0)
    np.testing.assert_allclose((s_a + s_b).cpu().numpy(), s_full.cpu().numpy(), rtol=2e-6)  # fp32 per-thread partials


def test_config4_effdet_d7_postprocess_topk_regime(lib, cuda):
    import torch
    from tfmv_b200 import synth
    from tfmv_b200.ai_models.efficientnet.utils.anchors import Anchors
    from tfmv_b200.ai_models.efficientnet.utils.nms import get_nms
    c = synth.EFFDET_CONFIGS["d7"]
    a = Anchors(c["min_level"], c["max_level"], c["image_size"], c["num_scales"], c["aspect_ratios"], c["anchor_scale"])
    batch = 2
    g = torch.Generator(device=cuda).manual_seed(4)
    rel = [torch.randn((batch,) + tuple(b.shape), device=cuda, generator=g) * 0.25 for b in a.boxes]
    cls = [torch.randn((batch,) + tuple(b.shape[:-1]) + (81,), device=cuda, generator=g) for b in a.boxes]
    assert sum(int(b.numel()) // 4 for b in a.boxes) == 441936
    dec = a.convert_outputs_boxes(rel)
    r = a.convert_outputs_batch(dec, cls, with_indices=True)
    cnt = r["count"].cpu().numpy()
    assert (cnt == 200).all()
    s = r["scores"]
    assert bool((s[:, 1:] <= s[:, :-1]).all()) and bool((s > 0.5).all())   # raw logit >= 1e-4 -> sigmoid > 0.5
    assert bool((r["classes_id"] != 0).all())
    for k in range(batch):                                               # idempotence of the NMS on its own output
        logit = torch.log(s[k] / (1 - s[k]))
        again = get_nms(r["boxes"][k], logit, 200, 0.5, 0.0001, "diou")
        assert again.cpu().tolist() == list(range(200))
    # the selected anchors really are class maxima: gather and check the score
    for k in range(batch):
        flat = torch.cat([c[k].reshape(-1, 81) for c in cls], 0)
        sel = r["sel_anchor"][k].long()
        mx, am = flat[sel].max(-1)
        assert torch.equal(am, r["classes_id"][k])
        assert torch.allclose(torch.sigmoid(mx), s[k], rtol=1e-5)


def test_config4_effdet_d7_image_matches_oracle(lib, cuda):
    """One full D7 image (441 936 anchors, ~436 k candidates: the multi-CTA pivot / pre-gather window path) through the
    oracle's convert_outputs_one (efficientnet/utils/anchors.py:161-202, nms.py:5-61) — ids bit-exact, boxes and
    scores equal.  The oracle needs a few seconds for its 200 NMS iterations over 436 k boxes."""
    import torch
    from oracle import effdet as oe
    from tfmv_b200 import synth
    from tfmv_b200.ai_models.efficientnet.utils.anchors import Anchors
    c = synth.EFFDET_CONFIGS["d7"]
    args = (c["min_level"], c["max_level"], c["image_size"], c["num_scales"], c["aspect_ratios"], c["anchor_scale"])
    a, o = Anchors(*args), oe.Anchors(*args)
    rng = np.random.default_rng(20261018 + 4)
    rel, cls = synth.effdet_heads(rng, 1, c["image_size"])
    d = lambda xs: [torch.from_numpy(x).to(cuda) for x in xs]
    dec = a.convert_outputs_boxes(d(rel))
    want_dec = o.convert_outputs_boxes(rel)
    for l in range(len(rel)):
        assert np.array_equal(dec[l].cpu().numpy(), want_dec[l])
    r = a.convert_outputs_batch(dec, d(cls), with_indices=True)
    want = o.convert_outputs_one_ex(0, want_dec, cls)
    k = int(r["count"][0])
    assert k == len(want["selected"]) == 200
    assert r["sel_idx"][0, :k].cpu().tolist() == want["selected"].tolist()          # NMS indices: bit-exact
    assert r["sel_anchor"][0, :k].cpu().tolist() == want["cand_anchor"][want["selected"]].tolist()
    assert r["classes_id"][0, :k].cpu().tolist() == want["classes_id"].tolist()
    assert np.array_equal(r["boxes"][0, :k].cpu().numpy(), want["boxes"])
    assert np.array_equal(r["scores"][0, :k].cpu().numpy(), want["scores"])

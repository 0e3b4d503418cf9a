"""GPU parity for the EfficientDet utilities against the oracle."""
import numpy as np
import pytest

from test_gpu_core import _t, assert_bits_equal

pytestmark = pytest.mark.gpu
F = np.float32
LOSS_RTOL = 1e-4
CFGS = {
    "tiny": dict(min_level=0, max_level=0, image_size=(10, 10), num_scales=3, aspect_ratios=[(1.0, 1.0)], anchor_scale=3.0),
    "small": dict(min_level=3, max_level=7, image_size=(128, 128), num_scales=3,
                  aspect_ratios=[(1.0, 1.0), (1.4, 0.7), (0.7, 1.4)], anchor_scale=4.0),
    "rect": dict(min_level=3, max_level=5, image_size=(128, 256), num_scales=2,
                 aspect_ratios=[(1.0, 1.0), (1.4, 0.7)], anchor_scale=[3.0, 4.0, 5.0]),
    "d0": dict(min_level=3, max_level=7, image_size=(512, 512), num_scales=3,
               aspect_ratios=[(1.0, 1.0), (1.4, 0.7), (0.7, 1.4)], anchor_scale=4.0),
}


def _pair(name):
    from oracle import effdet as oe
    from tfmv_b200.ai_models.efficientnet.utils.anchors import Anchors
    c = CFGS[name]
    args = (c["min_level"], c["max_level"], c["image_size"], c["num_scales"], c["aspect_ratios"], c["anchor_scale"])
    return Anchors(*args), oe.Anchors(*args)


@pytest.mark.parametrize("name", ["tiny", "small", "rect", "d0"])
def test_anchor_boxes_bit_exact(lib, cuda, name):
    a, o = _pair(name)
    assert len(a.boxes) == len(o.boxes) and a.get_anchors_per_location() == o.get_anchors_per_location()
    for g, w in zip(a.boxes, o.boxes):
        assert_bits_equal(g.cpu().numpy(), w)


@pytest.mark.parametrize("name,batch", [("small", 3), ("rect", 2), ("d0", 2)])
def test_decode_bit_exact(lib, cuda, name, batch):
    a, o = _pair(name)
    rng = np.random.default_rng(3)
    rel = [(rng.standard_normal((batch,) + b.shape, dtype=F) * F(0.25)) for b in o.boxes]
    rel[0][0, 0, 0, 0, 2] = 120.0  # exp overflow -> inf box, kept as is (anc:266)
    want = o.convert_outputs_boxes(rel)
    got = a.convert_outputs_boxes([_t(r, cuda) for r in rel])
    for g, w in zip(got, want):
        assert_bits_equal(g.cpu().numpy(), w)


def test_reference_anchor_fixture(lib, cuda):
    """tests/test_anchors.py:10-34 end to end: generate_targets -> convert_outputs_boxes -> convert_outputs_one."""
    a, o = _pair("tiny")
    boxes = np.array([[3, 3, 6, 6], [5, 5, 9, 9]], F)
    classes = np.array([1, 2])
    ob, oc, om = a.generate_targets(_t(boxes, cuda), classes, 3, iou_threshold=0.5)
    wb, wc, wm = o.generate_targets(boxes, classes, 3, iou_threshold=0.5)
    assert tuple(ob[0].shape) == (10, 10, 3, 4) and tuple(oc[0].shape) == (10, 10, 3, 3) and tuple(om[0].shape) == (10, 10, 3, 1)
    assert_bits_equal(ob[0].cpu().numpy(), wb[0])
    assert_bits_equal(oc[0].cpu().numpy(), wc[0])
    assert np.array_equal(om[0].cpu().numpy(), wm[0])
    dec = a.convert_outputs_boxes([ob[0][None]])
    b, c, s = a.convert_outputs_one(0, dec, [oc[0][None]])
    assert c.cpu().tolist() == [1, 2] and str(c.dtype) == "torch.int64"
    np.testing.assert_allclose(b.cpu().numpy(), [[3, 3, 6, 6], [5, 5, 9, 9]], atol=2e-5)
    np.testing.assert_allclose(s.cpu().numpy(), [0.7310586, 0.7310586], rtol=1e-6)


@pytest.mark.parametrize("name,batch,nmax", [("small", 3, 40), ("rect", 2, 10), ("d0", 2, 100)])
def test_generate_targets_bit_exact(lib, cuda, name, batch, nmax):
    from tfmv_b200 import synth
    a, o = _pair(name)
    rng = np.random.default_rng(20261018 + 3 + batch)
    ih, iw = CFGS[name]["image_size"]
    boxes, classes, off = synth.gt_batch(rng, batch, (iw, ih), max_boxes=nmax, order="yxyx")
    classes = (classes % 80 + 1).astype(np.int32)  # class 0 is background
    classes[0] = 200  # out of range -> all-zero one-hot row (tf.one_hot)
    C = 81
    gb, gc, gm = a.generate_targets_batch(_t(boxes, cuda), _t(classes, cuda), _t(off, cuda), C)
    npos = 0
    for b in range(batch):
        wb, wc, wm = o.generate_targets(boxes[off[b]:off[b + 1]], classes[off[b]:off[b + 1]], C)
        for l in range(len(wb)):
            assert_bits_equal(gb[l][b].cpu().numpy(), wb[l])
            assert_bits_equal(gc[l][b].cpu().numpy(), wc[l])
            assert np.array_equal(gm[l][b].cpu().numpy(), wm[l])
            npos += int(wm[l].sum())
    assert npos > 0


@pytest.mark.parametrize("thr", [0.0, 0.3, 0.9])
def test_generate_targets_other_thresholds(lib, cuda, thr):
    """iou_threshold = 0 matches every anchor (argmax of an all-zero IoU row is GT 0): the GT culling must be off there."""
    from tfmv_b200 import synth
    a, o = _pair("small")
    rng = np.random.default_rng(20261018 + 60)
    ih, iw = CFGS["small"]["image_size"]
    boxes, classes, off = synth.gt_batch(rng, 2, (iw, ih), max_boxes=12, order="yxyx")
    classes = (classes % 80 + 1).astype(np.int32)
    gb, gc, gm = a.generate_targets_batch(_t(boxes, cuda), _t(classes, cuda), _t(off, cuda), 81, iou_threshold=thr)
    for b in range(2):
        wb, wc, wm = o.generate_targets(boxes[off[b]:off[b + 1]], classes[off[b]:off[b + 1]], 81, iou_threshold=thr)
        for l in range(len(wb)):
            assert_bits_equal(gb[l][b].cpu().numpy(), wb[l])
            assert_bits_equal(gc[l][b].cpu().numpy(), wc[l])
            assert np.array_equal(gm[l][b].cpu().numpy(), wm[l])


@pytest.mark.parametrize("C", [1, 2, 4, 7])
def test_generate_targets_narrow_class_rows(lib, cuda, C):
    """One-hot rows narrower than a float4 (and odd widths) take the scalar / unaligned store paths."""
    from tfmv_b200 import synth
    a, o = _pair("small")
    rng = np.random.default_rng(20261018 + 40 + C)
    ih, iw = CFGS["small"]["image_size"]
    boxes, classes, off = synth.gt_batch(rng, 2, (iw, ih), max_boxes=20, order="yxyx")
    classes = (classes % (C + 1)).astype(np.int32)  # includes the out-of-range id C
    gb, gc, gm = a.generate_targets_batch(_t(boxes, cuda), _t(classes, cuda), _t(off, cuda), C)
    for b in range(2):
        wb, wc, wm = o.generate_targets(boxes[off[b]:off[b + 1]], classes[off[b]:off[b + 1]], C)
        for l in range(len(wb)):
            assert_bits_equal(gb[l][b].cpu().numpy(), wb[l])
            assert_bits_equal(gc[l][b].cpu().numpy(), wc[l])
            assert np.array_equal(gm[l][b].cpu().numpy(), wm[l])


@pytest.mark.parametrize("name,batch,iou_type", [("small", 3, "diou"), ("rect", 2, "ciou"), ("d0", 2, "diou"), ("small", 2, "giou")])
def test_convert_outputs_one_matches_oracle(lib, cuda, name, batch, iou_type):
    a, o = _pair(name)
    rng = np.random.default_rng(20261018 + 4 + batch)
    C = 81
    rel = [(rng.standard_normal((batch,) + b.shape, dtype=F) * F(0.25)) for b in o.boxes]
    cls = [rng.standard_normal((batch,) + b.shape[:-1] + (C,), dtype=F) for b in o.boxes]
    cls[0][0, 0, 0, :, :] = 0.5          # ties across classes -> argmax 0 -> background, dropped
    cls[0][0, 0, 1, :, 7] = 3.0          # exact score ties between anchors -> lower index first
    dec_w = o.convert_outputs_boxes(rel)
    dec = a.convert_outputs_boxes([_t(r, cuda) for r in rel])
    r = a.convert_outputs_batch(dec, [_t(c, cuda) for c in cls], iou_type=iou_type, with_indices=True)
    r = {k: v.cpu().numpy() for k, v in r.items()}
    for b in range(batch):
        w = o.convert_outputs_one_ex(b, dec_w, cls, iou_type=iou_type)
        k = int(r["count"][b])
        assert k == w["selected"].shape[0] and k > 0
        assert r["sel_idx"][b, :k].tolist() == w["selected"].tolist()
        assert r["sel_anchor"][b, :k].tolist() == w["cand_anchor"][w["selected"]].tolist()
        assert r["classes_id"][b, :k].tolist() == w["classes_id"].tolist()
        assert_bits_equal(r["boxes"][b, :k], w["boxes"])
        assert_bits_equal(r["scores"][b, :k], w["scores"])
    # reference signature, one image
    bx, ci, sc = a.convert_outputs_one(1, dec, [_t(c, cuda) for c in cls]) if iou_type == "diou" else (None, None, None)
    if bx is not None:
        w = o.convert_outputs_one_ex(1, dec_w, cls)
        assert ci.cpu().tolist() == w["classes_id"].tolist()
        assert_bits_equal(bx.cpu().numpy(), w["boxes"])


@pytest.mark.parametrize("name,batch", [("small", 4), ("d0", 2)])
def test_get_loss_matches_oracle(lib, cuda, name, batch):
    from oracle import effdet as oe
    from tfmv_b200 import synth
    from tfmv_b200.ai_models.efficientnet.efficientdet_net_train import get_loss
    from tfmv_b200.ai_models.losses.box_loss import BoxLoss
    from tfmv_b200.ai_models.losses.focal_loss import FocalLoss
    a, o = _pair(name)
    rng = np.random.default_rng(20261018 + 3)
    ih, iw = CFGS[name]["image_size"]
    C = 81
    boxes, classes, off = synth.gt_batch(rng, batch, (iw, ih), max_boxes=30, order="yxyx")
    classes = (classes % 80 + 1).astype(np.int32)
    per = [o.generate_targets(boxes[off[b]:off[b + 1]], classes[off[b]:off[b + 1]], C) for b in range(batch)]
    L = len(o.boxes)
    tb = [np.stack([p[0][l] for p in per], 0) for l in range(L)]
    tc = [np.stack([p[1][l] for p in per], 0) for l in range(L)]
    tm = [np.stack([p[2][l] for p in per], 0) for l in range(L)]
    pb = [(rng.standard_normal(t.shape, dtype=F) * F(0.25)) for t in tb]
    pc = [rng.standard_normal(t.shape, dtype=F) for t in tc]
    want, wparts, wnpos = oe.get_loss(tb, tc, tm, pb, pc, return_parts=True)
    d = lambda xs: [_t(x, cuda) for x in xs]
    got, parts, npos = get_loss(d(tb), d(tc), d(tm), d(pb), d(pc), return_parts=True)
    assert float(npos) == float(wnpos) > 1
    np.testing.assert_allclose(parts.cpu().numpy(), wparts, rtol=LOSS_RTOL, atol=1e-9)
    assert abs(float(got) - float(want)) <= LOSS_RTOL * abs(float(want))
    # stand-alone Keras-style losses
    fl = FocalLoss(0.25, 1.5)([float(wnpos), _t(tc[0], cuda)], _t(pc[0], cuda))
    assert abs(float(fl) - float(oe.focal_loss(wnpos, tc[0], pc[0]))) <= LOSS_RTOL * abs(float(oe.focal_loss(wnpos, tc[0], pc[0])))
    bl = BoxLoss()([float(wnpos), _t(tb[0], cuda)], _t(pb[0], cuda))
    assert abs(float(bl) - float(oe.box_loss(wnpos, tb[0], pb[0]))) <= LOSS_RTOL * abs(float(oe.box_loss(wnpos, tb[0], pb[0])))
    el = FocalLoss(0.25, 1.5).call([3.0, _t(tc[1], cuda)], _t(pc[1], cuda)).cpu().numpy()
    np.testing.assert_allclose(el, oe.focal_loss_elements(3.0, tc[1], pc[1]), rtol=2e-5, atol=1e-9)


@pytest.mark.parametrize("name,batch", [("small", 3), ("d0", 2), ("tiny", 2)])
def test_class_index_targets_match_one_hot_path(lib, cuda, name, batch):
    """SURVEY §8f N3 for EfficientDet: generate_targets with class ids instead of one-hot rows, and the focal loss
    rebuilt from them, equal the dense path (ids == argmax/zero-row of the one-hot, loss within 1e-4 of the oracle)."""
    import torch
    from oracle import effdet as oe
    from tfmv_b200 import synth
    from tfmv_b200.ai_models.efficientnet.efficientdet_net_train import get_loss
    a, o = _pair(name)
    rng = np.random.default_rng(20261018 + 90 + batch)
    ih, iw = CFGS[name]["image_size"]
    C = 81 if name != "tiny" else 3
    if name == "tiny":
        boxes = np.array([[3, 3, 6, 6], [5, 5, 9, 9], [2, 2, 5, 5]], F); classes = np.array([1, 2, 1], np.int32); off = np.array([0, 2, 3], np.int32)
    else:
        boxes, classes, off = synth.gt_batch(rng, batch, (iw, ih), max_boxes=40, order="yxyx")
        classes = (classes % 80 + 1).astype(np.int32)
        classes[0] = 200  # out of range -> all-zero one-hot row
    db, dc, do = _t(boxes, cuda), _t(classes, cuda), _t(off, cuda)
    gb, gc, gm = a.generate_targets_batch(db, dc, do, C)
    ib, ic, im = a.generate_targets_batch(db, dc, do, C, class_index=True)
    npos = 0
    for l in range(len(gb)):
        assert ic[l].dtype == torch.int32 and tuple(ic[l].shape) == tuple(gc[l].shape[:-1])
        assert torch.equal(ib[l], gb[l]) and torch.equal(im[l], gm[l])
        ids = ic[l].long()
        onehot = torch.zeros_like(gc[l])
        ok = (ids >= 0) & (ids < C)
        onehot[ok] = torch.nn.functional.one_hot(ids[ok], C).float()
        assert torch.equal(onehot, gc[l])
        npos += int(gm[l].sum())
    assert npos > 0
    pb = [torch.randn(t.shape, device=cuda, generator=torch.Generator(cuda).manual_seed(5 + l)) * 0.25 for l, t in enumerate(gb)]
    pc = [torch.randn(t.shape, device=cuda, generator=torch.Generator(cuda).manual_seed(50 + l)) for l, t in enumerate(gc)]
    dense, dparts, dn = get_loss(gb, gc, gm, pb, pc, return_parts=True)
    sparse, sparts, sn = get_loss(ib, ic, im, pb, pc, return_parts=True)
    assert float(dn) == float(sn)
    np.testing.assert_allclose(sparts.cpu().numpy(), dparts.cpu().numpy(), rtol=1e-6, atol=1e-12)
    want = oe.get_loss([t.cpu().numpy() for t in gb], [t.cpu().numpy() for t in gc], [t.cpu().numpy() for t in gm],
                       [t.cpu().numpy() for t in pb], [t.cpu().numpy() for t in pc])
    assert abs(float(sparse) - float(want)) <= LOSS_RTOL * abs(float(want))


def test_get_loss_gradient_matches_oracle(lib, cuda):
    """SURVEY §8f N1 for EfficientDet: d _get_loss / d class logits and box outputs vs the fp64 analytic oracle."""
    from oracle import effdet as oe
    from tfmv_b200 import synth
    from tfmv_b200.ai_models.efficientnet.efficientdet_net_train import get_loss_and_grad
    a, o = _pair("small")
    rng = np.random.default_rng(37)
    batch, C = 3, 81
    boxes, classes, off = synth.gt_batch(rng, batch, (128, 128), max_boxes=20, order="yxyx")
    classes = (classes % 80 + 1).astype(np.int32)
    per = [o.generate_targets(boxes[off[b]:off[b + 1]], classes[off[b]:off[b + 1]], C) for b in range(batch)]
    L = len(o.boxes)
    tb = [np.stack([p[0][l] for p in per], 0) for l in range(L)]
    tc = [np.stack([p[1][l] for p in per], 0) for l in range(L)]
    tm = [np.stack([p[2][l] for p in per], 0) for l in range(L)]
    pb = [(rng.standard_normal(t.shape, dtype=F) * F(0.25)) for t in tb]
    pc = [rng.standard_normal(t.shape, dtype=F) for t in tc]
    d = lambda xs: [_t(x, cuda) for x in xs]
    loss, gb, gc = get_loss_and_grad(d(tb), d(tc), d(tm), d(pb), d(pc))
    want = oe.get_loss(tb, tc, tm, pb, pc)
    assert abs(float(loss) - float(want)) <= LOSS_RTOL * abs(float(want))
    wb, wc = oe.get_loss_grad(tb, tc, tm, pb, pc)
    for l in range(L):
        np.testing.assert_allclose(gc[l].cpu().numpy(), wc[l], rtol=2e-4, atol=1e-12)
        np.testing.assert_allclose(gb[l].cpu().numpy(), wb[l], rtol=2e-4, atol=1e-12)


@pytest.mark.gpu
def test_gradient_with_class_index_targets_equals_one_hot(lib, cuda):
    """get_loss(with_grad=True) on sparse class-id targets (generate_targets_batch(class_index=True)): the gradient must
    equal the gradient against the one-hot rows those ids stand for, including ids outside [0, C) (all-zero row)."""
    import torch
    from tfmv_b200 import synth
    from tfmv_b200.ai_models.efficientnet.efficientdet_net_train import get_loss_and_grad
    a, o = _pair("small")
    rng = np.random.default_rng(53)
    batch, C = 3, 81
    boxes, classes, off = synth.gt_batch(rng, batch, (128, 128), max_boxes=20, order="yxyx")
    classes = (classes % 80 + 1).astype(np.int32)
    classes[::7] = 200  # tf.one_hot: out of range -> zeros
    d_b, d_c, d_o = _t(boxes, cuda), _t(classes, cuda), _t(off, cuda)
    tb, tc, tm = a.generate_targets_batch(d_b, d_c, d_o, C)
    ib, ic, im = a.generate_targets_batch(d_b, d_c, d_o, C, class_index=True)
    g = torch.Generator(device=cuda).manual_seed(5)
    pb = [torch.randn(t.shape, device=cuda, generator=g) * 0.25 for t in tb]
    pc = [torch.randn(t.shape, device=cuda, generator=g) for t in tc]
    loss_d, gb_d, gc_d = get_loss_and_grad(tb, tc, tm, pb, pc)
    loss_i, gb_i, gc_i = get_loss_and_grad(ib, ic, im, pb, pc)
    assert abs(float(loss_d) - float(loss_i)) <= 1e-5 * abs(float(loss_d))
    for l in range(len(tb)):
        assert torch.equal(gc_d[l], gc_i[l]), l     # same arithmetic on the same y values: identical bits
        assert torch.equal(gb_d[l], gb_i[l]), l
    with pytest.raises(ValueError):
        get_loss_and_grad(ib, [t[..., :1] for t in tc], im, pb, pc)


@pytest.mark.parametrize("name,batch,C,iou_type", [("small", 3, 81, "diou"), ("rect", 2, 7, "ciou"), ("d0", 2, 81, "diou"), ("small", 5, 6, "giou")])
def test_fused_eval_step_matches_oracle_and_separate_calls(lib, cuda, name, batch, C, iou_type):
    """test_step in one pass (b200_effdet_eval_step / b200_effdet_decode_postprocess): the class logits are read once for
    the focal loss and the argmax filter.  Against the oracle: decoded boxes, NMS indices / anchors / class ids / boxes /
    scores bit-exact, loss <= 1e-4; against the separate drop-in calls: identical outputs, loss equal to 1e-6."""
    import torch
    from oracle import effdet as oe
    from tfmv_b200 import synth
    from tfmv_b200.ai_models.efficientnet.efficientdet_net_train import get_loss
    a, o = _pair(name)
    rng = np.random.default_rng(600 + batch + C)
    hw = CFGS[name]["image_size"]
    boxes, classes, off = synth.gt_batch(rng, batch, (hw[1], hw[0]), max_boxes=20, order="yxyx")
    classes = (classes % (C - 1) + 1).astype(np.int32)
    per = [o.generate_targets(boxes[off[b]:off[b + 1]], classes[off[b]:off[b + 1]], C) for b in range(batch)]
    L = len(o.boxes)
    tb = [np.stack([p[0][l] for p in per], 0) for l in range(L)]
    tc = [np.stack([p[1][l] for p in per], 0) for l in range(L)]
    tm = [np.stack([p[2][l] for p in per], 0) for l in range(L)]
    rel = [(rng.standard_normal(t.shape, dtype=F) * F(0.25)) for t in tb]
    cls = [rng.standard_normal(t.shape, dtype=F) for t in tc]
    cls[0][0, 0, 0, :, :] = 0.5          # ties across classes -> argmax 0 -> background, dropped
    cls[0][0, 0, 1, :, C - 1] = 3.0      # exact score ties between anchors -> lower index first
    cls[0][1, 0, 0, 0, 0] = np.nan       # NaN in channel 0 sticks (serial scan from element 0): background
    cls[0][1, 0, 0, 1, 3] = np.nan       # NaN elsewhere is skipped by `v > m`
    rel[0][0, 0, 0, 0, 2] = 120.0        # exp overflow -> inf box (anc:266)
    d = lambda xs: [_t(x, cuda) for x in xs]
    loss, dec, r = a.eval_step(d(tb), d(tc), d(tm), d(rel), d(cls), iou_type=iou_type, with_indices=True)
    dec2, r2 = a.decode_and_postprocess(d(rel), d(cls), iou_type=iou_type, with_indices=True)
    dec_w = o.convert_outputs_boxes(rel)
    for l in range(L):
        assert_bits_equal(dec[l].cpu().numpy(), dec_w[l])
        assert_bits_equal(dec2[l].cpu().numpy(), dec_w[l])
    rn = {k: v.cpu().numpy() for k, v in r.items()}
    assert sorted(rn) == sorted(r2)
    for b in range(batch):
        # the image with planted NaN logits is compared with the separate calls only (np.argmax and the serial `v > m`
        # scan of the kernels differ on NaN; non-finite logits are outside the parity contract)
        w = None if np.isnan(cls[0][b]).any() else o.convert_outputs_one_ex(b, dec_w, cls, iou_type=iou_type)
        k = int(rn["count"][b])
        assert k == int(r2["count"][b])
        for key in ("sel_idx", "sel_anchor", "classes_id", "boxes", "scores"):
            assert np.array_equal(rn[key][b, :k], r2[key][b, :k].cpu().numpy(), equal_nan=True), (b, key)
        if w is None:
            continue
        assert k == w["selected"].shape[0] and k > 0
        assert rn["sel_idx"][b, :k].tolist() == w["selected"].tolist()
        assert rn["sel_anchor"][b, :k].tolist() == w["cand_anchor"][w["selected"]].tolist()
        assert rn["classes_id"][b, :k].tolist() == w["classes_id"].tolist()
        assert_bits_equal(rn["boxes"][b, :k], w["boxes"])
        assert_bits_equal(rn["scores"][b, :k], w["scores"])
    # the separate drop-in calls give the same post-processing result, NaN rows included
    rs = a.convert_outputs_batch(a.convert_outputs_boxes(d(rel)), d(cls), iou_type=iou_type, with_indices=True)
    for b in range(batch):
        k = int(rn["count"][b])
        assert k == int(rs["count"][b])
        for key in ("sel_idx", "sel_anchor", "classes_id", "boxes", "scores"):
            assert np.array_equal(rn[key][b, :k], rs[key][b, :k].cpu().numpy(), equal_nan=True), (b, key)
    cls_l = [np.nan_to_num(c, nan=0.0) for c in cls]           # the loss itself on NaN-free logits
    loss2, _, _ = a.eval_step(d(tb), d(tc), d(tm), d(rel), d(cls_l), iou_type=iou_type)
    want = float(oe.get_loss(tb, tc, tm, rel, cls_l))
    assert abs(float(loss2) - want) <= LOSS_RTOL * abs(want)
    sep = float(get_loss(d(tb), d(tc), d(tm), d(rel), d(cls_l)))
    assert abs(float(loss2) - sep) <= 2e-6 * abs(sep)


def test_ciou_v_custom_gradient_bitwise(lib, cuda):
    """_get_v + the gradient of its tf.custom_gradient (efficientnet/utils/iou.py:5-24) against the oracle, bit for bit
    (same detmath atan, same operation order)."""
    from oracle import effdet as oe
    from tfmv_b200.ai_models.efficientnet.utils.iou import _get_v
    rng = np.random.default_rng(91)
    n = 5000
    h1, w1, h2, w2 = [np.exp(rng.uniform(np.log(0.5), np.log(500.0), n)).astype(F) for _ in range(4)]
    h2[::97] = 0.0
    w1[::89] = 0.0
    h1[::83] = 0.0
    dv = rng.standard_normal(n).astype(F)
    v, gh, gw = _get_v(_t(h1, cuda), _t(w1, cuda), _t(h2, cuda), _t(w2, cuda), _t(dv, cuda))
    wv, wgh, wgw = oe.get_v_grad(h1, w1, h2, w2, dv)
    assert_bits_equal(v.cpu().numpy(), wv)
    assert_bits_equal(gh.cpu().numpy(), wgh)
    assert_bits_equal(gw.cpu().numpy(), wgw)
    assert_bits_equal(_get_v(_t(h1, cuda), _t(w1, cuda), _t(h2, cuda), _t(w2, cuda)).cpu().numpy(), wv)

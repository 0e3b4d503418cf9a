"""CPU tests pinning the oracle: the reference's two unit-test relations, hand-derived known answers for the
literal inputs the reference ships, and the TF op semantics the restatement relies on."""
import numpy as np
import pytest

from oracle import detmath as dm
from oracle import effdet as oe
from oracle import yolo as oy

F = np.float32


def test_grid_layout_relation():
    # yolo_v3/unit_test/grid_test.py:30-33 — meshgrid grid == keras-yolo3 tile grid, [...,0]=x, [...,1]=y
    for h, w in [(13, 13), (5, 7)]:
        g1, g2 = oy.grid_meshgrid(h, w), oy.grid_tile(h, w)
        assert g1.shape == (h, w, 1, 2) and (g1 == g2).all()
        assert g1[2, 3, 0, 0] == 3 and g1[2, 3, 0, 1] == 2


def test_getloss_copy_equals_yolov4loss():
    # yolo_v3/unit_test/loss_test.py:152-172 — exact fp32 equality on uniform-random tensors
    rng = np.random.default_rng(1)
    yt = [rng.random((2, g, g, 3, 85), dtype=F) for g in (2, 4, 8)]
    yp = [rng.random((2, g, g, 255), dtype=F) for g in (2, 4, 8)]
    anc = oy.load_anchors_order(oy.COCO_ANCHORS_FLAT.reshape(-1))
    a = oy.get_loss(yt, yp, (64, 64), anc, variant="unit_test_copy")
    b = oy.yolov4_loss(oy.COCO_ANCHORS_FLAT, 80, yt, yp)
    assert a == b and np.isfinite(a)


def test_iou_known_answers():
    # efficientnet/utils/iou.py:104-111: [10,10,30,30] vs [20,20,40,40]
    b1 = np.array([[10, 10, 30, 30]], F)
    b2 = np.array([[20, 20, 40, 40]], F)
    iou = F(100.0 / 700.0)
    assert oe.get_iou(b1, b2, "iou")[0] == iou
    assert abs(oe.get_iou(b1, b2, "diou")[0] - (100 / 700 - 200 / 1800)) < 1e-7
    assert oe.get_iou(b1, b2, "ciou")[0] == oe.get_iou(b1, b2, "diou")[0]  # v == 0 for squares
    assert abs(oe.get_iou(b1, b2, "giou")[0] - (100 / 700 - 200 / 900)) < 1e-7
    y = lambda t: oy.get_iou(b1[:, None, :], b2[None], t)[0, 0]
    assert y("iou") == iou
    assert abs(y("diou") - (100 / 700 - (1 / 9) ** 0.6)) < 2e-7
    assert abs(y("ciou") - (100 / 700 - 1 / 9)) < 1e-7
    with pytest.raises(AssertionError):
        oy.get_iou(b1[:, None, :], b2[None], "giou")  # tf_iou_utils.py:18 accepts only iou/diou/ciou


def test_yolo_iou_degenerate_semantics():
    # no clamp, plain divide: 0/0 -> NaN; c == 0 -> iou (tf_iou_utils.py:34,51)
    z = np.zeros((1, 1, 4), F)
    assert np.isnan(oy.get_iou(z, z, "iou")).all()
    p = np.array([[[1, 1, 1, 1]]], F)
    assert np.isnan(oy.get_iou(p, p, "diou")).all()


def test_nms_tie_and_nan_semantics():
    boxes = np.array([[0, 0, 1, 1], [0, 0, 1, 1], [2, 2, 3, 3], [2, 2, 3, 3.5]], F)
    scores = np.array([0.5, 0.5, 0.9, 0.9], F)
    # ties -> lower index first; duplicates suppressed
    assert oy.get_iou_nms(boxes, scores, 10, 0.5).tolist() == [2, 0]
    assert oy.get_iou_nms_by_classes(boxes, scores, np.array([0, 1, 0, 0]), 10, 0.5).tolist() == [2, 0, 1]
    # degenerate boxes: plain NMS drops NaN pairs, per-class NMS keeps them (tiu:98 vs :146)
    deg = np.zeros((3, 4), F)
    s = np.array([0.3, 0.2, 0.1], F)
    assert oy.get_iou_nms(deg, s, 10, 0.5).tolist() == [0]
    assert oy.get_iou_nms_by_classes(deg, s, np.zeros(3, np.int32), 10, 0.5).tolist() == [0, 1, 2]
    # max_output_size cap
    assert oy.get_iou_nms(boxes, scores, 1, 0.5).tolist() == [2]


def test_effdet_nms_score_threshold_stop():
    boxes = np.array([[0, 0, 10, 10], [20, 20, 30, 30], [40, 40, 50, 50]], F)
    scores = np.array([2.0, 0.00005, 1.0], F)
    assert oe.get_nms(boxes, scores, 200, 0.5, 0.0001, "diou").tolist() == [0, 2]


def test_anchor_fixture_known_answers():
    # tests/test_anchors.py:10-15: Anchors(0,0,(10,10),3,[(1,1)],3.0)
    a = oe.Anchors(0, 0, (10, 10), 3, [(1.0, 1.0)], 3.0)
    assert len(a.boxes) == 1 and a.boxes[0].shape == (10, 10, 3, 4)
    half = [3.0 * 2 ** (k / 3.0) / 2.0 for k in range(3)]  # 1.5, 1.8899, 2.3811
    for k in range(3):
        np.testing.assert_allclose(a.boxes[0][0, 0, k], [0.5 - half[k], 0.5 - half[k], 0.5 + half[k], 0.5 + half[k]], rtol=1e-6)
        np.testing.assert_allclose(a.boxes[0][9, 4, k], [9.5 - half[k], 4.5 - half[k], 9.5 + half[k], 4.5 + half[k]], rtol=1e-6)
    ob, oc, om = a.generate_targets(np.array([[3, 3, 6, 6], [5, 5, 9, 9]], F), np.array([1, 2]), 3, iou_threshold=0.5)
    # anchor (y=4,x=4,k=0) = [3,3,6,6] exactly -> IoU 1 with GT 0 -> class 1, zero regression -> masked by BoxLoss
    assert om[0][4, 4, 0, 0] and oc[0][4, 4, 0].tolist() == [0, 1, 0] and (ob[0][4, 4, 0] == 0).all()
    # unmatched anchors are one-hot class 0 (anchors.py:131-133)
    assert not om[0][0, 0, 0, 0] and oc[0][0, 0, 0].tolist() == [1, 0, 0]
    # round trip: decode(encode(gt)) returns the GT box for matched anchors
    dec = a.convert_outputs_boxes([ob[0][None]])
    m = om[0][..., 0]
    gt_for = np.where(oc[0][m][:, 1:2] == 1, np.array([[3, 3, 6, 6]], F), np.array([[5, 5, 9, 9]], F))
    np.testing.assert_allclose(dec[0][0][m], gt_for, atol=2e-5)
    r = a.convert_outputs_one_ex(0, dec, [oc[0][None]])
    assert r["classes_id"].tolist() == [1, 2]
    np.testing.assert_allclose(r["boxes"], [[3, 3, 6, 6], [5, 5, 9, 9]], atol=2e-5)
    np.testing.assert_allclose(r["scores"], [0.7310586, 0.7310586], rtol=1e-6)  # sigmoid(1)


def test_feat_sizes_and_anchor_counts():
    assert oe.get_feat_sizes((512, 512), 7)[3:] == [(64, 64), (32, 32), (16, 16), (8, 8), (4, 4)]
    a0 = oe.Anchors(3, 7, (512, 512), 3, [(1.0, 1.0), (1.4, 0.7), (0.7, 1.4)], 4.0)
    assert sum(b.shape[0] * b.shape[1] * b.shape[2] for b in a0.boxes) == 49104
    # aspect[1] scales x, aspect[0] scales y (anchors.py:66-67): aspect (1.4,0.7) is taller than wide
    b = a0.boxes[0][0, 0, 1]
    assert (b[2] - b[0]) > (b[3] - b[1])


def test_get_targets_quirks():
    anc = oy.load_anchors_order(oy.COCO_ANCHORS_FLAT.reshape(-1))
    assert anc[0].tolist() == [[116, 90], [156, 198], [373, 326]]  # largest first (LoadAnchors [2,1,0])
    boxes = np.array([[100, 120, 201, 260], [100, 120, 201, 260], [300, 40, 380, 90]], F)
    t = oy.get_targets(boxes, np.array([3, 5, 7]), anc, (416, 416), 80)
    assert [x.shape for x in t] == [(13, 13, 3, 85), (26, 26, 3, 85), (52, 52, 3, 85)]
    # normalised wh vs pixel anchors -> always the smallest anchor: flat idx 6 -> layer 2, anchor 0
    assert t[0].sum() == 0 and t[1].sum() == 0
    # two GTs in the same cell collide -> record zeroed; third survives with floor-div centre
    assert t[2][..., 4].sum() == 1.0
    cy, cx = int(np.floor(F(65.0 / 416) * 52)), int(np.floor(F(340.0 / 416) * 52))
    rec = t[2][cy, cx, 0]
    assert rec[4] == 1 and rec[5 + 7] == 1 and rec[0] == F(340.0) / F(416) and rec[2] == F(80) / F(416)
    # float floor-div of the centre: (100+201)//2 = 150, not 150.5
    t1 = oy.get_targets(boxes[:1], np.array([3]), anc, (416, 416), 80)
    y, x = np.argwhere(t1[2][..., 4] == 1)[0][:2]
    assert t1[2][y, x, 0, 0] == F(150.0) / F(416)


def test_losses_against_closed_forms():
    rng = np.random.default_rng(3)
    x = rng.normal(size=(4, 5, 5, 9, 7)).astype(F)
    y = (rng.random(x.shape) < 0.1).astype(F)
    e = oe.focal_loss_elements(3.0, y, x)
    p = 1 / (1 + np.exp(-x.astype(np.float64)))
    pt = y * p + (1 - y) * (1 - p)
    ref = (y * 0.25 + (1 - y) * 0.75) * (1 - pt) ** 1.5 * (-np.log(pt)) / 3.0
    np.testing.assert_allclose(e, ref, rtol=2e-5, atol=1e-9)
    assert abs(oe.focal_loss(3.0, y, x) - ref.mean()) < 1e-6 * abs(ref.mean()) + 1e-9
    t = rng.normal(size=(4, 5, 5, 9, 4)).astype(F) * (rng.random((4, 5, 5, 9, 4)) < 0.3)
    o = rng.normal(size=t.shape).astype(F) * F(0.2)
    err = np.abs(o.astype(np.float64) - t)
    hub = np.where(err <= 0.1, 0.5 * err ** 2, 0.1 * err - 0.005) * (t != 0)
    assert abs(oe.box_loss(7.0, t, o) - hub.sum() / 28.0) < 1e-5 * hub.sum() / 28.0


def test_detmath_against_numpy():
    rng = np.random.default_rng(5)
    x = (rng.normal(size=200000) * 6).astype(F)
    for f, g in [(dm.exp, np.exp), (dm.sigmoid, lambda v: 1 / (1 + np.exp(-v))), (dm.atan, np.arctan)]:
        ref = g(x.astype(np.float64))
        assert np.max(np.abs(f(x) - ref) / np.abs(ref)) < 4e-7
    xp = np.abs(x) + F(1e-6)
    assert np.max(np.abs(dm.log(xp) - np.log(xp.astype(np.float64))) / np.maximum(np.abs(np.log(xp.astype(np.float64))), 1e-30)) < 4e-7
    z = (rng.random(x.shape) < 0.5).astype(F)
    ref = np.maximum(x, 0) - x * z + np.log1p(np.exp(-np.abs(x.astype(np.float64))))
    np.testing.assert_allclose(dm.bce_logits(z, x), ref, rtol=1e-6, atol=1e-7)


def test_loss_gradient_matches_finite_differences():
    """The analytic backward of GetLoss (oracle.get_loss_grad) against fp64 central differences of the same loss."""
    from tfmv_b200 import synth
    rng = np.random.default_rng(23)
    image, batch = 64, 2
    anc = (synth.yolo_anchors().astype(F) / F(416)).astype(F)
    boxes, classes, off = synth.gt_batch(rng, batch, (image, image), max_boxes=6)
    per = [oy.get_targets(boxes[off[b]:off[b + 1]], classes[off[b]:off[b + 1]], anc, (image, image), 80) for b in range(batch)]
    y_true = [np.stack([p[l] for p in per], 0) for l in range(3)]
    y_pred = synth.yolo_heads(rng, batch, image)
    loss, grads = oy.get_loss_grad(y_true, y_pred, (image, image), anc, 0.5, "ciou")
    _, _, ign = oy.get_loss(y_true, y_pred, (image, image), anc, 0.5, "ciou", return_ignore=True)
    f = lambda yp: oy.loss_fp64_fixed_ignore(y_true, yp, (image, image), anc, ign)
    assert abs(f(y_pred) - float(loss)) < 1e-4 * abs(float(loss))
    eps = 1e-4
    checked = 0
    for l in range(3):
        g = grads[l].reshape(y_pred[l].shape)
        flat_idx = list(rng.integers(0, y_pred[l].size, 12))
        obj_rec = np.argwhere(y_true[l][..., 4] > 0)
        for r in obj_rec[:2]:   # make sure object records (xy, wh, class channels) are covered
            for c in (0, 1, 2, 3, 4, 5 + int(np.argmax(y_true[l][tuple(r)][5:])), 9):
                flat_idx.append(np.ravel_multi_index((r[0], r[1], r[2], r[3] * 85 + c), y_pred[l].shape))
        for fi in flat_idx:
            idx = np.unravel_index(int(fi), y_pred[l].shape)
            yp_p = [a.astype(np.float64) for a in y_pred]
            yp_m = [a.astype(np.float64) for a in y_pred]
            yp_p[l][idx] += eps
            yp_m[l][idx] -= eps
            fd = (f(yp_p) - f(yp_m)) / (2 * eps)
            assert abs(fd - float(g[idx])) <= 2e-4 * max(abs(fd), 1e-3), (l, idx, fd, float(g[idx]))
            checked += 1
    assert checked > 40


def test_effdet_loss_gradient_matches_finite_differences():
    rng = np.random.default_rng(31)
    shapes = [(2, 4, 4, 9), (2, 2, 2, 9)]
    tc = [(rng.random(s + (11,)) < 0.08).astype(F) for s in shapes]
    pc = [rng.standard_normal(s + (11,)).astype(F) for s in shapes]
    tb = [(rng.standard_normal(s + (4,)) * (rng.random(s + (4,)) < 0.3)).astype(F) for s in shapes]
    pb = [(rng.standard_normal(s + (4,)) * 0.2).astype(F) for s in shapes]
    tm = [rng.random(s + (1,)) < 0.2 for s in shapes]
    assert abs(oe.loss_fp64(tb, tc, tm, pb, pc) - float(oe.get_loss(tb, tc, tm, pb, pc))) < 1e-5 * abs(float(oe.get_loss(tb, tc, tm, pb, pc)))
    gb, gc = oe.get_loss_grad(tb, tc, tm, pb, pc)
    eps = 1e-5
    for l in range(2):
        for arr, grad, which in ((pc, gc, "c"), (pb, gb, "b")):
            for fi in rng.integers(0, arr[l].size, 10):
                idx = np.unravel_index(int(fi), arr[l].shape)
                ap = [a.astype(np.float64) for a in arr]
                am = [a.astype(np.float64) for a in arr]
                ap[l][idx] += eps
                am[l][idx] -= eps
                fp = oe.loss_fp64(tb, tc, tm, ap if which == "b" else pb, ap if which == "c" else pc)
                fm = oe.loss_fp64(tb, tc, tm, am if which == "b" else pb, am if which == "c" else pc)
                fd = (fp - fm) / (2 * eps)
                assert abs(fd - grad[l][idx]) <= 1e-4 * max(abs(fd), 1e-6), (which, l, idx, fd, grad[l][idx])

#!/usr/bin/env python
"""Runs the reference's OWN source files for the hot path (imported unmodified from /root/reference/AIServer) on seeded
inputs under the NumPy stand-in for TensorFlow in tests/golden/fake_tf, and stores inputs + outputs in
tests/golden/ref_emulated.npz.  tests/test_reference_emulated.py then holds the oracle to these outputs: integer
results exactly, floating-point results to a few ulp (the stand-in uses libm transcendentals, the oracle detmath).

This pins the oracle's restatement of the reference's control flow, operation order, broadcasting and index
conventions to the reference's actual code.  It does not pin TensorFlow's kernels (see fake_tf/tensorflow/__init__.py).

    python tests/golden/make_golden_emulated.py        (needs /root/reference; not run on the GPU box)
"""
import os
import sys

import numpy as np

import importlib.util

HERE = os.path.dirname(os.path.abspath(__file__))
# a real TensorFlow, when the image has one, takes precedence over the stand-in (B200_FORCE_TF_STAND_IN=1 overrides)
REAL_TF = importlib.util.find_spec("tensorflow") is not None and not os.environ.get("B200_FORCE_TF_STAND_IN")
if not REAL_TF:
    sys.path.insert(0, os.path.join(HERE, "fake_tf"))
sys.path.insert(0, "/root/reference/AIServer")
F = np.float32


def main():
    from ai_api.ai_models.utils import tf_iou_utils as riou
    from ai_api.ai_models.utils import tf_yolo_utils as ryolo
    from ai_api.ai_models.efficientnet.utils import iou as reiou
    from ai_api.ai_models.efficientnet.utils import nms as renms
    from ai_api.ai_models.efficientnet.utils.anchors import Anchors as RAnchors
    from ai_api.ai_models.losses.focal_loss import FocalLoss as RFocal
    from ai_api.ai_models.losses.box_loss import BoxLoss as RBox
    from ai_api.ai_models.losses.yolo_loss import Yolov4Loss as RYolov4Loss
    rng = np.random.default_rng(20261018 + 77)
    out = {}

    def boxes_xyxy(n, scale=1.0):
        c = rng.uniform(0.1, 0.9, (n, 2)); wh = np.exp(rng.uniform(np.log(0.03), np.log(0.5), (n, 2)))
        return (np.concatenate([c - wh / 2, c + wh / 2], 1) * scale).astype(F)

    # ---- tf_iou_utils -------------------------------------------------------------------------------------
    b1, b2 = boxes_xyxy(40), boxes_xyxy(17)
    b2[3] = b1[5]                        # identical pair
    b2[4, 2:] = b2[4, :2]                # zero-area box (0/0 cases)
    out["iou_b1"], out["iou_b2"] = b1, b2
    for t in ("iou", "diou", "ciou"):
        out["iou_" + t] = riou.GetIOU(b1[:, None, :], b2[None, :, :], t)
    nb = boxes_xyxy(300)
    nb[50:80] = nb[20:50] + rng.normal(0, 0.004, (30, 4)).astype(F)   # near duplicates: suppression
    ns = rng.random(300).astype(F)
    ns[100:110] = ns[90]                 # exact score ties
    nc = rng.integers(0, 4, 300).astype(np.int32)
    out["nms_boxes"], out["nms_scores"], out["nms_classes"] = nb, ns, nc
    for t in ("iou", "diou", "ciou"):
        out["nms_plain_" + t] = riou.GetIOUNMS(nb, ns, 500, 0.5, t)
        out["nms_class_" + t] = riou.GetIOUNMSByClasses(nb, ns, nc, 500, 0.45, t)
    out["nms_plain_cap"] = riou.GetIOUNMS(nb, ns, 25, 0.5, "iou")

    # ---- tf_yolo_utils: GetBoxes / GetNMSBoxes / GetLoss ----------------------------------------------------
    image, C = 96, 6
    anchors = np.array([116, 90, 156, 198, 373, 326, 30, 61, 62, 45, 59, 119, 10, 13, 16, 30, 33, 23], F).reshape(3, 3, 2) * F(96 / 416)
    grids = (3, 6, 12)
    heads = [(rng.standard_normal((1, g, g, 3 * (5 + C))) * 1.5).astype(F) for g in grids]
    heads[2][0, 0, 0, 2] = 95.0          # exp overflow -> inf -> 0 -> dropped
    out["y_heads0"], out["y_heads1"], out["y_heads2"], out["y_anchors"] = heads[0], heads[1], heads[2], anchors
    r = ryolo.GetNMSBoxes(heads[0], heads[1], heads[2], anchors, np.array([image, image], np.int32), C, 0.5, 0.3, 0.5, "diou")
    for k, v in zip(("boxes", "classes_id", "scores", "classes", "confidence"), r):
        out["y_nms_" + k] = np.asarray(v)
    gb = ryolo.GetBoxes(heads[1].reshape(1, 6, 6, 3, 5 + C), (anchors[1] / F(image)), C)
    out["y_getboxes_boxes"], out["y_getboxes_conf"], out["y_getboxes_classes"] = [np.asarray(v) for v in gb]
    B = 2
    y_true, y_pred = [], []
    for g in grids:
        t = np.zeros((B, g, g, 3, 5 + C), F)
        n_obj = max(2, g * g // 6)
        idx = rng.integers(0, [B, g, g, 3], (n_obj, 4))
        for b, yy, xx, a in idx:
            t[b, yy, xx, a, 0:2] = [(xx + rng.random()) / g, (yy + rng.random()) / g]
            t[b, yy, xx, a, 2:4] = np.exp(rng.uniform(np.log(0.05), np.log(0.6), 2))
            t[b, yy, xx, a, 4] = 1
            t[b, yy, xx, a, 5 + rng.integers(0, C)] = 1
        p = rng.standard_normal((B, g, g, 3 * (5 + C))).astype(F)
        # some predictions on top of their targets so that the ignore mask has zeros
        pr = p.reshape(B, g, g, 3, 5 + C)
        for b, yy, xx, a in idx[: n_obj // 2]:
            a2 = (a + 1) % 3
            fx, fy = t[b, yy, xx, a, 0] * g - xx, t[b, yy, xx, a, 1] * g - yy
            fx, fy = min(max(fx, 1e-3), 1 - 1e-3), min(max(fy, 1e-3), 1 - 1e-3)
            pr[b, yy, xx, a2, 0:2] = [np.log(fx / (1 - fx)), np.log(fy / (1 - fy))]
        y_true.append(t); y_pred.append(p)
    for l in range(3):
        for b, yy, xx, a in np.argwhere(y_true[l][..., 4] > 0):
            a2 = (a + 1) % 3
            y_pred[l].reshape(B, grids[l], grids[l], 3, 5 + C)[b, yy, xx, a2, 2:4] = np.log(y_true[l][b, yy, xx, a, 2:4] * image / anchors[l][a2])
    for l in range(3):
        out["yl_true%d" % l], out["yl_pred%d" % l] = y_true[l], y_pred[l]
    for t in ("iou", "diou", "ciou"):
        out["yl_loss_" + t] = np.asarray(ryolo.GetLoss(y_true, y_pred, np.array([image, image], np.int32), anchors, 0.5, t), dtype=F)
    flat9 = (np.array([10, 13, 16, 30, 33, 23, 30, 61, 62, 45, 59, 119, 116, 90, 156, 198, 373, 326], F).reshape(9, 2) * F(96 / 416)).astype(F)
    out["yl_yolov4loss"] = np.asarray(RYolov4Loss(flat9, C).call(y_true, [p.copy() for p in y_pred]), dtype=F)
    out["yl_anchors9"] = flat9

    # ---- efficientnet/utils: iou, nms, Anchors ---------------------------------------------------------------
    e1, e2 = boxes_xyxy(30, 100.0), boxes_xyxy(11, 100.0)
    e2[2] = e1[4]; e2[3, 2:] = e2[3, :2]
    out["e_b1"], out["e_b2"] = e1, e2
    for t in ("iou", "giou", "diou", "ciou"):
        out["e_iou_" + t] = reiou.get_iou(e1[:, None, :], e2[None, :, :], t)
    eb = boxes_xyxy(400, 128.0); eb[100:140] = eb[40:80] + rng.normal(0, 0.5, (40, 4)).astype(F)
    es = (rng.standard_normal(400) * 2).astype(F); es[10:14] = es[9]
    out["e_nms_boxes"], out["e_nms_scores"] = eb, es
    for t in ("iou", "giou", "diou", "ciou"):
        out["e_nms_" + t] = renms.get_nms(eb, es, 200, 0.5, 0.0001, t)
    out["e_nms_cap"] = renms.get_nms(eb, es, 7, 0.5, float("-inf"), "diou")
    cfg = dict(min_level=3, max_level=5, image_size=(64, 96), num_scales=2, aspect_ratios=[(1.0, 1.0), (1.4, 0.7)], anchor_scale=3.0)
    ra = RAnchors(**cfg)
    L = len(ra.boxes)
    for l in range(L):
        out["ea_boxes%d" % l] = np.asarray(ra.boxes[l])
    gtb = np.array([[5, 8, 40, 60], [20, 30, 60, 90], [2, 2, 12, 14], [30, 10, 50, 34]], F)
    gtc = np.array([1, 3, 7, 2], np.int32)     # 7 is out of range for classes_num = 5
    out["ea_gt_boxes"], out["ea_gt_classes"] = gtb, gtc
    tb, tc, tm = ra.generate_targets(gtb, gtc, 5, 0.5)
    for l in range(L):
        out["ea_tb%d" % l], out["ea_tc%d" % l], out["ea_tm%d" % l] = np.asarray(tb[l]), np.asarray(tc[l]), np.asarray(tm[l])
    Bc, Cc = 2, 5
    rel = [(rng.standard_normal((Bc,) + np.asarray(b).shape) * 0.3).astype(F) for b in ra.boxes]
    cls = [rng.standard_normal((Bc,) + np.asarray(b).shape[:-1] + (Cc,)).astype(F) for b in ra.boxes]
    dec = ra.convert_outputs_boxes(rel)
    for l in range(L):
        out["ea_rel%d" % l], out["ea_cls%d" % l], out["ea_dec%d" % l] = rel[l], cls[l], np.asarray(dec[l])
    for b in range(Bc):
        bx, ci, sc = ra.convert_outputs_one(b, dec, cls)
        out["ea_one%d_boxes" % b], out["ea_one%d_ids" % b], out["ea_one%d_scores" % b] = np.asarray(bx), np.asarray(ci), np.asarray(sc)
    # ---- losses ----------------------------------------------------------------------------------------------
    yt = (rng.random((2, 4, 4, 6, 5)) < 0.1).astype(F); yp = rng.standard_normal((2, 4, 4, 6, 5)).astype(F)
    out["fl_true"], out["fl_pred"] = yt, yp
    out["fl_elem"] = np.asarray(RFocal(0.25, 1.5).call((F(3.0), yt), yp))
    out["fl_mean"] = np.asarray(RFocal(0.25, 1.5)((F(3.0), yt), yp), dtype=F)
    bt = (rng.standard_normal((2, 4, 4, 6, 4)) * (rng.random((2, 4, 4, 6, 1)) < 0.2)).astype(F)
    bp = (rng.standard_normal((2, 4, 4, 6, 4)) * 0.2).astype(F)
    out["bl_true"], out["bl_pred"] = bt, bp
    out["bl_loss"] = np.asarray(RBox(0.1).call((F(7.0), bt), bp), dtype=F)
    # ---- datasets/coco_dataset.py: DataGenerator.GetTargets (the class is not constructed: its __init__ loads files) ----
    import types
    for m in ["matplotlib", "matplotlib.colors", "cv2", "PIL", "PIL.Image", "PIL.ImageFilter"]:
        if m not in sys.modules:
            try:
                __import__(m)
            except Exception:
                mod = types.ModuleType(m)
                mod.__getattr__ = lambda name: None
                sys.modules[m] = mod
    from ai_api.ai_models.datasets import coco_dataset as rcd
    coco = np.array([116, 90, 156, 198, 373, 326, 30, 61, 62, 45, 59, 119, 10, 13, 16, 30, 33, 23], F).reshape(3, 3, 2)
    for tag, anc in (("px", coco), ("norm", coco / F(416))):   # pixel anchors (the reference's call) and normalised ones
        gen = object.__new__(rcd.DataGenerator)
        gen.anchors_wh, gen.image_wh, gen.classes_num = anc, np.array([416, 416], F), 20   # TF would convert these to float32 tensors
        gen.layers_hw = [[13, 13], [26, 26], [52, 52]]
        n = 30
        c = rng.uniform(20, 396, (n, 2)); wh = np.exp(rng.uniform(np.log(8), np.log(300), (n, 2)))
        gtb = np.clip(np.concatenate([c - wh / 2, c + wh / 2], 1), 0, 416).astype(F)
        gtb[7] = gtb[3]; gtb[8] = gtb[3]                       # triple collision: the record is zeroed
        gtc = rng.integers(0, 20, n).astype(np.int32)
        _, tg = gen.GetTargets("img", gtc, gtb)
        out["gt_%s_boxes" % tag], out["gt_%s_classes" % tag], out["gt_%s_anchors" % tag] = gtb, gtc, anc
        for l in range(3):
            out["gt_%s_t%d" % (tag, l)] = np.asarray(tg[l])
    # ---- efficientdet_net_train._get_loss (edt:41-52), ClassFocalLoss, yolo_v4/model.py GetGroudTruth ------------
    for m in ["matplotlib.pyplot", "tensorflow_addons"]:
        if m not in sys.modules:
            try:
                __import__(m)
            except Exception:
                mod = types.ModuleType(m)
                mod.__getattr__ = lambda name: None
                sys.modules[m] = mod
    from ai_api.ai_models.efficientnet import efficientdet_net_train as redt
    from ai_api.ai_models.losses.class_loss import ClassFocalLoss as RClassFocal
    from ai_api.ai_models.yolo_v4 import model as rm4
    shapes = [(2, 8, 8, 6), (2, 4, 4, 6), (2, 2, 2, 6)]
    glb = [(rng.standard_normal(sh + (4,)) * (rng.random(sh + (1,)) < 0.15)).astype(F) for sh in shapes]
    glm = [(np.abs(t).sum(-1, keepdims=True) > 0) for t in glb]
    glc = [np.eye(5, dtype=F)[np.where(m[..., 0], rng.integers(1, 5, m.shape[:-1]), 0)] for m in glm]
    gpb = [(rng.standard_normal(sh + (4,)) * 0.3).astype(F) for sh in shapes]
    gpc = [rng.standard_normal(sh + (5,)).astype(F) for sh in shapes]
    stub = types.SimpleNamespace(box_loss=RBox(), focal_loss=RFocal(0.25, 1.5, label_smoothing=0.0), _reg_l2_loss=lambda wd: F(0.0))
    out["gl_loss"] = np.asarray(redt.EfficientDetNetTrain._get_loss(stub, glb, glc, glm, gpb, gpc), dtype=F)
    for l in range(3):
        out["gl_tb%d" % l], out["gl_tc%d" % l], out["gl_tm%d" % l], out["gl_pb%d" % l], out["gl_pc%d" % l] = glb[l], glc[l], glm[l], gpb[l], gpc[l]
    glm2 = [m.copy() for m in glm]; glm2[2][:] = False           # a level without positives: divide_no_nan -> 0
    out["cfl_loss"] = np.asarray(RClassFocal(0.25, 1.5).call(glc, (gpc, glm2)), dtype=F)
    gstub = types.SimpleNamespace(classes_num=6)
    for l in range(3):
        out["ggt%d" % l] = np.asarray(rm4.YoloV4Model.GetGroudTruth(gstub, y_true[l]))
    # ---- the reference's own tests/test_anchors.py main(), unmodified: capture what it prints --------------------
    import tensorflow as tf
    captured = {}

    def capture(*a, **k):
        if len(a) == 2 and isinstance(a[0], str):
            captured[a[0].rstrip(":")] = np.asarray(a[1])
    tf.print = capture
    sys.path.insert(0, "/root/reference/AIServer/ai_api/ai_models/tests")
    import test_anchors as rtest
    rtest.tf.print = capture
    rtest.main()
    for k in ("convert_boxes", "convert_classes_id", "convert_scores"):
        out["rt_" + k] = captured[k]
    # ---- BASELINE config 1 through the reference's GetNMSBoxes: 416x416, 80 classes, ~5.3 k candidates, cap 500 ----
    sys.path.insert(0, HERE)
    import emulated_inputs as ei
    h416 = ei.yolo_416_heads()
    for t, thr in (("iou", 0.5), ("diou", 0.45)):
        r = ryolo.GetNMSBoxes(h416[0], h416[1], h416[2], ei.COCO_ANCHORS, np.array([416, 416], np.int32), 80, 0.5, 0.3, thr, t)
        out["c1_%s_boxes" % t], out["c1_%s_ids" % t], out["c1_%s_scores" % t] = np.asarray(r[0]), np.asarray(r[1]), np.asarray(r[2])
        out["c1_%s_conf" % t] = np.asarray(r[4])
        out["c1_%s_classes_rowsum" % t] = np.asarray(r[3]).sum(-1, dtype=np.float64)   # the (500,80) block as a checksum
    # ---- BASELINE config 3 per image through the reference's Anchors: D0 pyramid (49 104 anchors), cap 200 ---------
    rd0 = RAnchors(**ei.D0)
    shapes0 = [tuple(np.asarray(b).shape[:3]) for b in rd0.boxes]
    rel0, cls0 = ei.effdet_d0_heads(shapes0)
    dec0 = rd0.convert_outputs_boxes(rel0)
    bx, ci, sc = rd0.convert_outputs_one(0, dec0, cls0)
    out["c3_boxes"], out["c3_ids"], out["c3_scores"] = np.asarray(bx), np.asarray(ci), np.asarray(sc)
    out["c3_anchor_checksum"] = np.array([np.asarray(b, dtype=np.float64).sum() for b in rd0.boxes])
    # ---- efficientnet/utils/iou.py:5-24 _get_v and the gradient its tf.custom_gradient returns (drawn last: every array
    # above keeps its values) ----------------------------------------------------------------------------------------
    if not REAL_TF:
        import tensorflow as tfs
        vh1, vw1, vh2, vw2 = [np.exp(rng.uniform(np.log(2.0), np.log(200.0), 64)).astype(F) for _ in range(4)]
        vh2[5] = 0.0; vw1[9] = 0.0          # divide_no_nan branches
        vdv = rng.standard_normal(64).astype(F)
        out["cv_h1"], out["cv_w1"], out["cv_h2"], out["cv_w2"], out["cv_dv"] = vh1, vw1, vh2, vw2, vdv
        out["cv_v"] = np.asarray(reiou._get_v(vh1, vw1, vh2, vw2), dtype=F)
        (gdh, gdw), _ = tfs.custom_gradient.last_grad(vdv, [])
        out["cv_gdh"], out["cv_gdw"] = np.asarray(gdh, dtype=F), np.asarray(gdw, dtype=F)
    path = os.environ.get("B200_EMULATED_OUT") or os.path.join(HERE, "ref_emulated.npz")   # the risk study writes elsewhere
    np.savez_compressed(path, **out)
    print("wrote %s: %d arrays (%s)" % (path, len(out), "real TensorFlow" if REAL_TF else "NumPy stand-in for TensorFlow"))


if __name__ == "__main__":
    main()

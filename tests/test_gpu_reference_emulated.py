"""GPU parity against the reference's OWN source files run under the NumPy stand-in for TensorFlow
(tests/golden/ref_emulated.npz, see tests/test_reference_emulated.py): the CUDA path, through the drop-in shims, on the
same inputs — indices / class ids / masks / targets identical, values to a few ulp (libm vs detmath)."""
import os

import numpy as np
import pytest

from test_gpu_core import _t

pytestmark = pytest.mark.gpu
F = np.float32
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_emulated.npz"))


def close(a, b, rtol=3e-6, atol=3e-6):
    a = a.cpu().numpy() if hasattr(a, "cpu") else np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    np.testing.assert_allclose(np.nan_to_num(a, nan=0.0), np.nan_to_num(b, nan=0.0), rtol=rtol, atol=atol)


def test_yolo_iou_nms_decode(lib, cuda):
    from tfmv_b200.ai_models.utils.tf_iou_utils import GetIOU, GetIOUNMS, GetIOUNMSByClasses
    from tfmv_b200.ai_models.utils.tf_yolo_utils import GetNMSBoxes
    for t in ("iou", "diou", "ciou"):
        close(GetIOU(_t(G["iou_b1"][:, None, :], cuda), _t(G["iou_b2"][None, :, :], cuda), t), G["iou_" + t])
        b, s, c = _t(G["nms_boxes"], cuda), _t(G["nms_scores"], cuda), _t(G["nms_classes"], cuda)
        assert GetIOUNMS(b, s, 500, 0.5, t).cpu().tolist() == G["nms_plain_" + t].tolist()
        assert GetIOUNMSByClasses(b, s, c, 500, 0.45, t).cpu().tolist() == G["nms_class_" + t].tolist()
    r = GetNMSBoxes(_t(G["y_heads0"], cuda), _t(G["y_heads1"], cuda), _t(G["y_heads2"], cuda), G["y_anchors"], (96, 96), 6, 0.5, 0.3,
                    0.5, "diou")
    assert r[1].cpu().tolist() == G["y_nms_classes_id"].tolist()
    close(r[0], G["y_nms_boxes"]); close(r[2], G["y_nms_scores"]); close(r[3], G["y_nms_classes"]); close(r[4], G["y_nms_confidence"])


def test_yolo_targets_and_losses(lib, cuda):
    from tfmv_b200.ai_models.datasets.coco_dataset import DataGenerator
    from tfmv_b200.ai_models.losses.yolo_loss import Yolov4Loss
    from tfmv_b200.ai_models.utils.tf_yolo_utils import GetGroudTruth, GetLoss
    for tag in ("px", "norm"):
        gen = DataGenerator(20, G["gt_%s_anchors" % tag], (416, 416))
        _, tg = gen.GetTargets("img", G["gt_%s_classes" % tag], _t(G["gt_%s_boxes" % tag], cuda))
        for l in range(3):
            assert np.array_equal(tg[l].cpu().numpy(), G["gt_%s_t%d" % (tag, l)])
    y_true = [_t(G["yl_true%d" % l], cuda) for l in range(3)]
    y_pred = [_t(G["yl_pred%d" % l], cuda) for l in range(3)]
    for t in ("iou", "diou", "ciou"):
        want = float(G["yl_loss_" + t])
        assert abs(float(GetLoss(y_true, y_pred, (96, 96), G["y_anchors"], 0.5, t)) - want) <= 1e-4 * abs(want)
    want = float(G["yl_yolov4loss"])
    assert abs(float(Yolov4Loss(G["yl_anchors9"], 6)(y_true, y_pred)) - want) <= 1e-4 * abs(want)
    for l in range(3):
        assert np.array_equal(GetGroudTruth(y_true[l]).cpu().numpy(), G["ggt%d" % l])


def test_effdet(lib, cuda):
    from tfmv_b200.ai_models.efficientnet.efficientdet_net_train import get_loss
    from tfmv_b200.ai_models.efficientnet.utils.anchors import Anchors
    from tfmv_b200.ai_models.efficientnet.utils.iou import get_iou
    from tfmv_b200.ai_models.efficientnet.utils.nms import get_nms
    from tfmv_b200.ai_models.losses.box_loss import BoxLoss
    from tfmv_b200.ai_models.losses.focal_loss import FocalLoss
    for t in ("iou", "giou", "diou", "ciou"):
        close(get_iou(_t(G["e_b1"][:, None, :], cuda), _t(G["e_b2"][None, :, :], cuda), t), G["e_iou_" + t])
        assert get_nms(_t(G["e_nms_boxes"], cuda), _t(G["e_nms_scores"], cuda), 200, 0.5, 0.0001, t).cpu().tolist() == G["e_nms_" + t].tolist()
    a = Anchors(3, 5, (64, 96), 2, [(1.0, 1.0), (1.4, 0.7)], 3.0)
    L = len(a.boxes)
    for l in range(L):
        assert np.array_equal(a.boxes[l].cpu().numpy(), G["ea_boxes%d" % l])
    tb, tc, tm = a.generate_targets(_t(G["ea_gt_boxes"], cuda), G["ea_gt_classes"], 5, 0.5)
    for l in range(L):
        close(tb[l], G["ea_tb%d" % l])
        assert np.array_equal(tc[l].cpu().numpy(), G["ea_tc%d" % l]) and np.array_equal(tm[l].cpu().numpy(), G["ea_tm%d" % l])
    rel = [_t(G["ea_rel%d" % l], cuda) for l in range(L)]
    cls = [_t(G["ea_cls%d" % l], cuda) for l in range(L)]
    dec = a.convert_outputs_boxes(rel)
    for l in range(L):
        close(dec[l], G["ea_dec%d" % l], rtol=3e-6, atol=3e-5)
    for b in range(2):
        bx, ci, sc = a.convert_outputs_one(b, dec, cls)
        assert ci.cpu().tolist() == G["ea_one%d_ids" % b].tolist()
        close(bx, G["ea_one%d_boxes" % b], rtol=3e-6, atol=3e-5); close(sc, G["ea_one%d_scores" % b])
    d = lambda k: [_t(G["%s%d" % (k, l)], cuda) for l in range(3)]
    want = float(G["gl_loss"])
    assert abs(float(get_loss(d("gl_tb"), d("gl_tc"), d("gl_tm"), d("gl_pb"), d("gl_pc"))) - want) <= 1e-4 * abs(want)
    want = float(G["fl_mean"])
    assert abs(float(FocalLoss(0.25, 1.5)([3.0, _t(G["fl_true"], cuda)], _t(G["fl_pred"], cuda))) - want) <= 1e-4 * abs(want)
    want = float(G["bl_loss"])
    assert abs(float(BoxLoss(0.1)([7.0, _t(G["bl_true"], cuda)], _t(G["bl_pred"], cuda))) - want) <= 1e-4 * abs(want)


def test_baseline_config1_against_the_reference_get_nms_boxes(lib, cuda):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import emulated_inputs as ei
    from tfmv_b200.ai_models.utils.tf_yolo_utils import GetNMSBoxes
    heads = [_t(h, cuda) for h in ei.yolo_416_heads()]
    for t, thr in (("iou", 0.5), ("diou", 0.45)):
        r = GetNMSBoxes(heads[0], heads[1], heads[2], ei.COCO_ANCHORS, (416, 416), 80, 0.5, 0.3, thr, t)
        assert r[1].cpu().tolist() == G["c1_%s_ids" % t].tolist() and len(r[1]) == 500
        close(r[0], G["c1_%s_boxes" % t]); close(r[2], G["c1_%s_scores" % t]); close(r[4], G["c1_%s_conf" % t])
        np.testing.assert_allclose(r[3].cpu().numpy().sum(-1, dtype=np.float64), G["c1_%s_classes_rowsum" % t], rtol=1e-5)


def test_baseline_config3_image_against_the_reference_anchors(lib, cuda):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import emulated_inputs as ei
    from tfmv_b200.ai_models.efficientnet.utils.anchors import Anchors
    a = Anchors(**ei.D0)
    rel, cls = ei.effdet_d0_heads([tuple(b.shape[:3]) for b in a.boxes])
    dec = a.convert_outputs_boxes([_t(r, cuda) for r in rel])
    bx, ci, sc = a.convert_outputs_one(0, dec, [_t(c, cuda) for c in cls])
    assert ci.cpu().tolist() == G["c3_ids"].tolist() and len(ci) == 200
    close(bx, G["c3_boxes"], rtol=3e-6, atol=1e-4); close(sc, G["c3_scores"])

"""mAP row.  The golden vectors in tests/golden/map_ref.npz were produced by the REFERENCE's own
NumPy code (tests/golden/make_golden_map.py imports /root/reference/.../utils/mAP.py), so this row is pinned by the
reference itself: the oracle restatement and the CUDA kernel are both checked against them."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "map_ref.npz")


def _cases():
    g = np.load(GOLD)
    return [(g["gt%d" % k], g["pr%d" % k], int(g["classes%d" % k]), float(g["map%d" % k])) for k in range(int(g["count"]))]


def test_oracle_matches_reference_generated_golden():
    from oracle import map as om
    cases = _cases()
    assert abs(cases[0][3] - 2.0 / 9.0) < 1e-15  # the literal example of mAP.py:130-142, first image
    for gt, pr, c, want in cases:
        assert abs(om.get_map_one(gt, pr, c, 0.5) - want) < 1e-12


@pytest.mark.gpu
def test_gpu_map_matches_reference_generated_golden(lib, cuda):
    import torch
    from tfmv_b200.ai_models.utils.mAP import Get_mAP_batch, Get_mAP_one
    cases = _cases()
    for gt, pr, c, want in cases:
        got = float(Get_mAP_one(torch.from_numpy(gt).to(cuda), torch.from_numpy(pr).to(cuda), c, 0.5))
        assert abs(got - want) < 1e-12, (gt.shape, pr.shape, got, want)
    # batched call over the cases that share class_num 80
    sel = [k for k, cs in enumerate(cases) if cs[2] == 80]
    go = np.cumsum([0] + [cases[k][0].shape[0] for k in sel]).astype(np.int32)
    po = np.cumsum([0] + [cases[k][1].shape[0] for k in sel]).astype(np.int32)
    out = Get_mAP_batch(np.concatenate([cases[k][0] for k in sel]), go, np.concatenate([cases[k][1] for k in sel]), po, 80)
    for j, k in enumerate(sel):
        assert abs(float(out[j]) - cases[k][3]) < 1e-12


@pytest.mark.gpu
def test_gpu_map_matches_oracle_on_nms_output(lib, cuda):
    """End of test_step: NMS output + GetGroudTruth rows -> mAP (yolo_v4/model.py:357-377)."""
    import torch
    from oracle import map as om
    from tfmv_b200 import synth
    from tfmv_b200.ai_models.datasets.coco_dataset import DataGenerator
    from tfmv_b200.ai_models.utils.mAP import Get_mAP_one
    from tfmv_b200.ai_models.utils.tf_yolo_utils import GetGroudTruth, GetNMSBoxes
    rng = np.random.default_rng(41)
    anc = synth.yolo_anchors().astype(np.float32)
    heads = [torch.from_numpy(h).to(cuda) for h in synth.yolo_heads_trained_like(rng, 1, 416)]
    boxes, classes, off = synth.gt_batch(rng, 1, (416, 416), max_boxes=30)
    gen = DataGenerator(80, anc / np.float32(416), (416, 416))
    y = gen.GetTargetsBatch(classes, boxes, off)
    sb, sc, ss, _, _ = GetNMSBoxes(*heads, anchors_wh=anc, image_wh=(416, 416), classes_num=80, confidence_thresh=0.5,
                                   scores_thresh=0.3, iou_thresh=0.5, iou_type='diou')
    pred = torch.cat([sb, sc.float()[:, None], ss[:, None]], dim=-1)
    gt = torch.cat([GetGroudTruth(t) for t in y], dim=0)
    got = float(Get_mAP_one(gt, pred, 80, 0.5))
    want = om.get_map_one(gt.cpu().numpy(), pred.cpu().numpy(), 80, 0.5)
    assert abs(got - want) < 1e-12


def _degenerate_case():
    # a zero-area ground-truth box and an identical zero-area prediction give IoU = 0/0 = NaN; np.argmax picks the first
    # NaN as the maximum and `NaN >= thresh` is False, so that ground-truth box claims nothing (mAP.py:53-56)
    gt = np.array([[10, 10, 10, 10, 0], [0, 0, 20, 20, 0]], dtype=np.float64)
    pr = np.array([[0, 0, 20, 20, 0, 0.9], [10, 10, 10, 10, 0, 0.8], [1, 1, 19, 19, 0, 0.7]], dtype=np.float64)
    return gt, pr


def test_oracle_nan_iou_follows_np_argmax():
    from oracle import map as om
    gt, pr = _degenerate_case()
    tp, n_gt = om.get_tpfp_one(gt, pr, 0, 0.5)
    # GT 0: IoU column = [0, NaN, 0] -> argmax = 1 (first NaN) -> no true positive; GT 1 claims prediction 0
    assert n_gt == 2 and tp[:, 0].tolist() == [1.0, 0.0, 0.0]


@pytest.mark.gpu
def test_gpu_map_nan_iou_follows_np_argmax(lib, cuda):
    import torch
    from oracle import map as om
    from tfmv_b200.ai_models.utils.mAP import Get_mAP_one
    gt, pr = _degenerate_case()
    # put the NaN pair first in the prediction list as well: the kernel must freeze on the first NaN it meets
    for order in ([0, 1, 2], [1, 0, 2], [2, 1, 0]):
        p = pr[order]
        got = float(Get_mAP_one(torch.from_numpy(gt).to(cuda), torch.from_numpy(p).to(cuda), 3, 0.5))
        assert abs(got - om.get_map_one(gt, p, 3, 0.5)) < 1e-12, order

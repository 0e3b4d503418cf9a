"""NumPy oracle for the mAP evaluation that follows NMS in test_step — TEST INFRASTRUCTURE ONLY.

Restates utils/mAP.py:3-125 (Get_TPFP, Get_AP, Get_mAP, Get_mAP_one) of the reference in float64.  This is the one
piece of the path that is pure NumPy in the reference, so it IS pinned by the reference itself:
tests/golden/make_golden_map.py imports /root/reference/.../utils/mAP.py (with the removed alias np.float restored)
and freezes its outputs in tests/golden/map_ref.npz; tests/test_map.py checks this restatement against them.
Quirks kept: the names precision/recall are swapped inside Get_AP (mAP.py:88-89), a prediction is a true positive when
it is the arg-max IoU prediction of some ground-truth box with IoU >= thresh (several boxes may claim the same one),
classes without predictions or without ground truth contribute AP = 0, the mean is over class_num.
"""
import numpy as np


def get_tpfp_one(groud_truth, prediction, class_id, thresh=0.5):
    gt = np.asarray(groud_truth, dtype=np.float64).reshape(-1, 5)
    pr = np.asarray(prediction, dtype=np.float64).reshape(-1, 6)
    gt = gt[gt[:, 4] == class_id]
    pr = pr[pr[:, 4] == class_id]
    n_gt = gt.shape[0]
    if n_gt == 0 or pr.shape[0] == 0:
        return np.zeros((0, 2)), n_gt
    g = gt[None, :, :]
    p = pr[:, None, :]
    imin = np.maximum(g[..., 0:2], p[..., 0:2])
    imax = np.minimum(g[..., 2:4], p[..., 2:4])
    iwh = np.maximum(imax - imin, 0.0)
    inter = iwh[..., 0] * iwh[..., 1]
    ga = (g[..., 2] - g[..., 0]) * (g[..., 3] - g[..., 1])
    pa = (p[..., 2] - p[..., 0]) * (p[..., 3] - p[..., 1])
    with np.errstate(all="ignore"):
        iou = inter / (ga + pa - inter)
    tp = np.zeros((pr.shape[0],))
    best = np.argmax(iou, axis=0)
    for i in range(best.shape[0]):
        if iou[best[i], i] >= thresh:
            tp[best[i]] = 1
    return np.stack([tp, pr[:, 5]], axis=-1), n_gt


def get_ap(tp, n_gt):
    tp = tp[np.argsort(tp[:, 1], kind="stable")[::-1], :] if tp.shape[0] else tp
    prec, rec = [], []
    s = 0.0
    for i in range(tp.shape[0]):
        if tp[i][0] == 1:
            s += 1.0
        prec.append(s / (i + 1))
        rec.append(s / n_gt)
    mrec = np.concatenate(([0.0], prec, [1.0]))
    mpre = np.concatenate(([0.0], rec, [0.0]))
    for i in range(mpre.size - 1, 0, -1):
        mpre[i - 1] = np.maximum(mpre[i - 1], mpre[i])
    i = np.where(mrec[1:] != mrec[:-1])[0]
    return float(np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1]))


def get_map_one(groud_truth, prediction, class_num, thresh=0.5):
    total = 0.0
    for c in range(int(class_num)):
        tp, n_gt = get_tpfp_one(groud_truth, prediction, c, thresh)
        total += get_ap(tp, n_gt)
    return total / class_num

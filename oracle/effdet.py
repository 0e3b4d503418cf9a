"""NumPy oracle for the EfficientDet part of the path — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates efficientnet/utils/{anchors,iou,nms,get_feat_sizes}.py, losses/{focal_loss,box_loss,class_loss}.py
and efficientnet/efficientdet_net_train.py:41-52 in fp32 NumPy, op by op.  Paths are relative to
/root/reference/AIServer/ai_api/ai_models/.
"""
import math

import numpy as np

from . import detmath as dm

F = np.float32
EPSILON = 1e-8  # anchors.py:10


def get_feat_sizes(image_size, max_level):
    """efficientnet/utils/get_feat_sizes.py:4-20."""
    fs = (int(image_size[0]), int(image_size[1]))
    sizes = [fs]
    for _ in range(1, max_level + 1):
        fs = ((fs[0] - 1) // 2 + 1, (fs[1] - 1) // 2 + 1)
        sizes.append(fs)
    return sizes


def _dnn(x, y):
    """tf.math.divide_no_nan: 0 where y == 0."""
    with np.errstate(all="ignore"):
        q = x / y
    return np.where(y == F(0), F(0), q).astype(F)


def _tf_range_f32(start, limit, delta):
    """tf.range on float32 (TF 2.3-2.8 CPU kernel: size = ceil((limit-start)/delta), val += delta)."""
    start, limit, delta = F(start), F(limit), F(delta)
    size = int(math.ceil(abs((float(limit) - float(start)) / float(delta))))
    out = np.empty((size,), dtype=F)
    v = start
    for i in range(size):
        out[i] = v
        v = F(v + delta)
    return out


def get_iou(boxes1, boxes2, iou_type="iou"):
    """efficientnet/utils/iou.py:26-100 get_iou on yxyx boxes (NaN-free family)."""
    b1 = np.asarray(boxes1, F)
    b2 = np.asarray(boxes2, F)
    b1_ymin, b1_xmin, b1_ymax, b1_xmax = b1[..., 0], b1[..., 1], b1[..., 2], b1[..., 3]
    b2_ymin, b2_xmin, b2_ymax, b2_xmax = b2[..., 0], b2[..., 1], b2[..., 2], b2[..., 3]
    zero = F(0)
    b1_w = np.maximum(zero, b1_xmax - b1_xmin)
    b1_h = np.maximum(zero, b1_ymax - b1_ymin)
    b2_w = np.maximum(zero, b2_xmax - b2_xmin)
    b2_h = np.maximum(zero, b2_ymax - b2_ymin)
    b1_area = b1_w * b1_h
    b2_area = b2_w * b2_h
    i_ymin = np.maximum(b1_ymin, b2_ymin)
    i_xmin = np.maximum(b1_xmin, b2_xmin)
    i_ymax = np.minimum(b1_ymax, b2_ymax)
    i_xmax = np.minimum(b1_xmax, b2_xmax)
    i_w = np.maximum(zero, i_xmax - i_xmin)
    i_h = np.maximum(zero, i_ymax - i_ymin)
    inter = i_w * i_h
    union = (b1_area + b2_area) - inter
    iou_v = _dnn(inter, union)
    if iou_type == "iou":
        return iou_v
    e_ymin = np.minimum(b1_ymin, b2_ymin)
    e_xmin = np.minimum(b1_xmin, b2_xmin)
    e_ymax = np.maximum(b1_ymax, b2_ymax)
    e_xmax = np.maximum(b1_xmax, b2_xmax)
    assert iou_type in ("giou", "diou", "ciou")
    if iou_type == "giou":
        e_w = np.maximum(zero, e_xmax - e_xmin)
        e_h = np.maximum(zero, e_ymax - e_ymin)
        e_area = e_w * e_h
        return iou_v - _dnn(e_area - union, e_area)
    c1y, c1x = (b1_ymin + b1_ymax) / F(2), (b1_xmin + b1_xmax) / F(2)
    c2y, c2x = (b2_ymin + b2_ymax) / F(2), (b2_xmin + b2_xmax) / F(2)
    # tf.linalg.norm(axis=-1) = sqrt(reduce_sum(x*x))
    dy, dx = c2y - c1y, c2x - c1x
    euclid = np.sqrt(dy * dy + dx * dx)
    ey, ex = e_ymax - e_ymin, e_xmax - e_xmin
    diag = np.sqrt(ey * ey + ex * ex)
    # `x**2` on a tensor is tf.pow(x, 2) == x*x in fp32
    diou_v = iou_v - _dnn(euclid * euclid, diag * diag)
    if iou_type == "diou":
        return diou_v
    arct = dm.atan(_dnn(b1_w, b1_h)) - dm.atan(_dnn(b2_w, b2_h))
    q = arct / F(math.pi)
    v = F(4) * (q * q)
    alpha = _dnn(v, (F(1) - iou_v) + v)
    return diou_v - alpha * v


def get_v_grad(b1_height, b1_width, b2_height, b2_width, dv):
    """efficientnet/utils/iou.py:5-24 _get_v with its custom gradient: (v, gdh, gdw) for the second box's (height, width)."""
    h1, w1, h2, w2, dv = [np.asarray(x, F) for x in (b1_height, b1_width, b2_height, b2_width, dv)]
    arct = (dm.atan(_dnn(w1, h1)) - dm.atan(_dnn(w2, h2))).astype(F)
    v = (F(4) * np.square(arct / F(math.pi))).astype(F)
    gdw = (dv * F(8) * arct * h2 / F(math.pi ** 2)).astype(F)
    gdh = (-dv * F(8) * arct * w2 / F(math.pi ** 2)).astype(F)
    return v, gdh, gdw


def get_nms(boxes, scores, max_output_size, iou_threshold=0.5, score_threshold=float("-inf"), iou_type="diou"):
    """efficientnet/utils/nms.py:5-61 get_nms: class-agnostic greedy, stop at first top score < score_threshold."""
    boxes = np.asarray(boxes, F).reshape(-1, 4)
    scores = np.asarray(scores, F).reshape(-1)
    order = np.argsort(-scores, kind="stable").astype(np.int32)
    cur = boxes[order]
    out = []
    while len(out) < max_output_size and cur.shape[0] > 0:
        if scores[order[0]] < F(score_threshold):
            break
        out.append(int(order[0]))
        if cur.shape[0] == 1:
            break
        m = get_iou(cur[0:1], cur[1:], iou_type=iou_type)
        keep = m < F(iou_threshold)
        cur = cur[1:][keep]
        order = order[1:][keep]
    return np.asarray(out, dtype=np.int32)


class Anchors(object):
    """efficientnet/utils/anchors.py:12-274."""

    def __init__(self, min_level, max_level, image_size, num_scales, aspect_ratios, anchor_scale):
        self.min_level = min_level
        self.max_level = max_level
        self.image_size = image_size
        self.num_scales = num_scales
        self.aspect_ratios = aspect_ratios
        if isinstance(anchor_scale, (list, tuple)):
            assert len(anchor_scale) == max_level - min_level + 1
            self.anchor_scales = list(anchor_scale)
        else:
            self.anchor_scales = [anchor_scale] * (max_level - min_level + 1)
        self.feat_sizes = get_feat_sizes(image_size, max_level)
        self.boxes = self._generate_boxes()

    def _generate_boxes(self):
        """anchors.py:46-84; note aspect[1] scales x and aspect[0] scales y (anc:66-67)."""
        fs = self.feat_sizes
        out = []
        for level in range(self.min_level, self.max_level + 1):
            per = []
            for octave in range(self.num_scales):
                for aspect in self.aspect_ratios:
                    stride = (fs[0][0] / float(fs[level][0]), fs[0][1] / float(fs[level][1]))
                    octave_scale = octave / float(self.num_scales)
                    a_scale = self.anchor_scales[level - self.min_level]
                    base_x = a_scale * stride[1] * 2 ** octave_scale
                    base_y = a_scale * stride[0] * 2 ** octave_scale
                    hx = base_x * aspect[1] / 2.0
                    hy = base_y * aspect[0] / 2.0
                    x = _tf_range_f32(stride[1] / 2, self.image_size[1], stride[1])
                    y = _tf_range_f32(stride[0] / 2, self.image_size[0], stride[0])
                    xv, yv = np.meshgrid(x, y)
                    yv = yv[..., None]
                    xv = xv[..., None]
                    b = np.concatenate([yv - F(hy), xv - F(hx), yv + F(hy), xv + F(hx)], axis=-1).astype(F)
                    per.append(b[..., None, :])
            out.append(np.concatenate(per, axis=-2))
        return out

    def get_anchors_per_location(self):
        return self.num_scales * len(self.aspect_ratios)

    @staticmethod
    def _cs(boxes):
        yc = (boxes[..., 2] + boxes[..., 0]) / F(2.0)
        xc = (boxes[..., 3] + boxes[..., 1]) / F(2.0)
        h = boxes[..., 2] - boxes[..., 0]
        w = boxes[..., 3] - boxes[..., 1]
        return yc[..., None], xc[..., None], h[..., None], w[..., None]

    def _boxes_encoder(self, anchors, boxes):
        """anchors.py:219-243."""
        yca, xca, ha, wa = self._cs(anchors)
        yc, xc, h, w = self._cs(boxes)
        ha = np.maximum(F(EPSILON), ha)
        wa = np.maximum(F(EPSILON), wa)
        h = np.maximum(F(EPSILON), h)
        w = np.maximum(F(EPSILON), w)
        tx = (xc - xca) / wa
        ty = (yc - yca) / ha
        tw = dm.log(w / wa)
        th = dm.log(h / ha)
        return np.concatenate([ty, tx, th, tw], axis=-1)

    def _boxes_decoder(self, anchors, rel):
        """anchors.py:245-274."""
        yca, xca, ha, wa = self._cs(anchors)
        ty, tx, th, tw = rel[..., 0:1], rel[..., 1:2], rel[..., 2:3], rel[..., 3:4]
        with np.errstate(all="ignore"):
            w = dm.exp(tw) * wa
            h = dm.exp(th) * ha
        yc = ty * ha + yca
        xc = tx * wa + xca
        return np.concatenate([yc - h / F(2.0), xc - w / F(2.0), yc + h / F(2.0), xc + w / F(2.0)], axis=-1)

    def generate_targets(self, boxes, classes, classes_num, iou_threshold=0.5):
        """anchors.py:91-138: anchor -> best GT (first max), mask = max >= thr, encode, one-hot (bg = class 0)."""
        boxes = np.asarray(boxes, F).reshape(-1, 4)
        classes = np.asarray(classes).reshape(-1, 1)
        ob, oc, om = [], [], []
        for anc in self.boxes:
            iou = get_iou(anc[..., None, :], boxes)
            idx = np.argmax(iou, axis=-1)
            mx = np.max(iou, axis=-1)
            mask = (mx >= F(iou_threshold))[..., None]
            bl = boxes[idx]
            cl = classes[idx]
            bl = self._boxes_encoder(anc, bl)
            bl = np.where(mask, bl, F(0))
            cl = np.where(mask, cl, np.zeros_like(cl))
            ci = cl[..., 0].astype(np.int32)
            onehot = (ci[..., None] == np.arange(classes_num)).astype(F)  # out-of-range -> all zeros
            ob.append(bl.astype(F))
            oc.append(onehot)
            om.append(mask)
        return tuple(ob), tuple(oc), tuple(om)

    def convert_outputs_boxes(self, outputs_boxes):
        """anchors.py:141-158."""
        return tuple(self._boxes_decoder(self.boxes[l], np.asarray(outputs_boxes[l], F)) for l in range(len(self.boxes)))

    def convert_outputs_one_ex(self, batch_index, outputs_boxes, outputs_classes, max_output_size=200,
                               iou_threshold=0.5, score_threshold=0.0001, iou_type="diou"):
        """anchors.py:161-202 plus the candidate list (for GPU compaction-order checks)."""
        nb, nc, ns, na = [], [], [], []
        base = 0
        for l in range(len(outputs_classes)):
            c = np.asarray(outputs_classes[l][batch_index], F)
            cid = np.argmax(c, axis=-1)
            cs = np.max(c, axis=-1)
            b = np.asarray(outputs_boxes[l][batch_index], F)
            m = cid != 0
            nb.append(b[m])
            nc.append(cid[m].astype(np.int64))
            ns.append(cs[m])
            na.append(np.flatnonzero(m.reshape(-1)) + base)
            base += m.size
        nb, nc, ns, na = np.concatenate(nb, 0), np.concatenate(nc, 0), np.concatenate(ns, 0), np.concatenate(na, 0)
        sel = get_nms(nb, ns, max_output_size=max_output_size, iou_threshold=iou_threshold,
                      score_threshold=score_threshold, iou_type=iou_type)
        return dict(boxes=nb[sel], classes_id=nc[sel], scores=dm.sigmoid(ns[sel]), selected=sel, cand_anchor=na,
                    cand_boxes=nb, cand_scores=ns, cand_classes_id=nc)

    def convert_outputs_one(self, batch_index, outputs_boxes, outputs_classes):
        r = self.convert_outputs_one_ex(batch_index, outputs_boxes, outputs_classes)
        return r["boxes"], r["classes_id"], r["scores"]


# --------------------------------------------------------------------------------------------
def focal_loss_elements(normalizer, y_true, y_pred, alpha=0.25, gamma=1.5, label_smoothing=0.0):
    """losses/focal_loss.py:26-52 FocalLoss.call: per-element alpha*mod*ce/normalizer."""
    y_true = np.asarray(y_true, F)
    y_pred = np.asarray(y_pred, F)
    a = F(alpha)
    p = dm.sigmoid(y_pred)
    p_t = (y_true * p) + ((F(1) - y_true) * (F(1) - p))
    af = y_true * a + (F(1) - y_true) * (F(1) - a)
    q = F(1.0) - p_t
    mod = dm.pow15(q) if gamma == 1.5 else dm.pow(q, gamma)
    yt = y_true * F(1.0 - label_smoothing) + F(0.5 * label_smoothing)
    ce = dm.bce_logits(yt, y_pred)
    return af * mod * ce / F(normalizer)


def focal_loss(normalizer, y_true, y_pred, **kw):
    """FocalLoss.__call__: Keras default reduction SUM_OVER_BATCH_SIZE == mean over every element."""
    e = focal_loss_elements(normalizer, y_true, y_pred, **kw)
    return F(np.sum(e.astype(np.float64)) / e.size)


def box_loss(num_positives, box_targets, box_outputs, delta=0.1):
    """losses/box_loss.py:17-29 BoxLoss.call: Huber(delta, NONE) on last-dim-1 tensors, masked by target != 0."""
    t = np.asarray(box_targets, F)
    o = np.asarray(box_outputs, F)
    normalizer = F(F(num_positives) * F(4.0))
    mask = (t != F(0.0)).astype(F)
    err = o - t  # keras Huber: error = y_pred - y_true
    ab = np.abs(err)
    d = F(delta)
    quad = F(0.5) * np.square(err)
    lin = d * ab - F(0.5) * np.square(d)
    hub = np.where(ab <= d, quad, lin).astype(F)  # mean over the size-1 last axis is the identity
    s = F(np.sum((hub * mask).astype(np.float64)))
    return F(s / normalizer)


def get_loss(y_true_boxes, y_true_classes, y_true_masks, y_pred_boxes, y_pred_classes, alpha=0.25, gamma=1.5,
             return_parts=False):
    """efficientnet/efficientdet_net_train.py:41-52 _get_loss without the L2-regularisation term
    (which is over model weights and stays in TF): sum_l (50*box_l + focal_l), num_pos = sum(mask) + 1."""
    npos = F(0.0)
    for m in y_true_masks:
        npos = F(npos + F(np.sum(np.asarray(m).astype(np.float64))))
    npos = F(npos + F(1.0))
    loss = F(0.0)
    parts = []
    for l in range(len(y_true_boxes)):
        lb = box_loss(npos, y_true_boxes[l], y_pred_boxes[l])
        lc = focal_loss(npos, y_true_classes[l], y_pred_classes[l], alpha=alpha, gamma=gamma)
        parts.append((lb, lc))
        loss = F(loss + F(F(lb * F(50.0)) + lc))
    if return_parts:
        return loss, np.asarray(parts, F), npos
    return loss


def class_focal_loss(class_targets, class_outputs, masks, alpha=0.25, gamma=1.5, label_smoothing=0.0):
    """losses/class_loss.py:25-60 ClassFocalLoss.call (demo-only variant): normalizer_l = sum(mask_l)/B, summed."""
    total = F(0.0)
    for t, o, m in zip(class_targets, class_outputs, masks):
        m = np.asarray(m).astype(F)
        normalizer = F(F(np.sum(m.astype(np.float64))) / F(m.shape[0]))
        e = focal_loss_elements(1.0, t, o, alpha=alpha, gamma=gamma, label_smoothing=label_smoothing)
        e = _dnn(e, np.full_like(e, normalizer))
        total = F(total + F(np.sum(e.astype(np.float64))))
    return total


def get_loss_grad(y_true_boxes, y_true_classes, y_true_masks, y_pred_boxes, y_pred_classes, alpha=0.25, gamma=1.5, delta=0.1):
    """d get_loss / d y_pred_boxes, d / d y_pred_classes in float64 (what tf.GradientTape derives; num_positives is
    a constant of the targets).  Pinned to central differences in tests/test_oracle_pins.py."""
    npos = 1.0 + sum(float(np.sum(np.asarray(m).astype(np.float64))) for m in y_true_masks)
    gb, gc = [], []
    for l in range(len(y_true_boxes)):
        y = np.asarray(y_true_classes[l], np.float64)
        x = np.asarray(y_pred_classes[l], np.float64)
        p = 1.0 / (1.0 + np.exp(-x))
        p_t = y * p + (1 - y) * (1 - p)
        af = y * alpha + (1 - y) * (1 - alpha)
        q = 1 - p_t
        ce = np.maximum(x, 0) - x * y + np.log1p(np.exp(-np.abs(x)))
        dq = -(2 * y - 1) * p * (1 - p)
        g = af * (gamma * np.power(q, gamma - 1) * dq * ce + np.power(q, gamma) * (p - y))
        gc.append(g / (npos * x.size))
        t = np.asarray(y_true_boxes[l], np.float64)
        o = np.asarray(y_pred_boxes[l], np.float64)
        e = o - t
        gb.append(np.where(t != 0, np.where(np.abs(e) <= delta, e, delta * np.sign(e)), 0.0) * 50.0 / (4.0 * npos))
    return gb, gc


def loss_fp64(y_true_boxes, y_true_classes, y_true_masks, y_pred_boxes, y_pred_classes, alpha=0.25, gamma=1.5, delta=0.1):
    npos = 1.0 + sum(float(np.sum(np.asarray(m).astype(np.float64))) for m in y_true_masks)
    total = 0.0
    for l in range(len(y_true_boxes)):
        y = np.asarray(y_true_classes[l], np.float64)
        x = np.asarray(y_pred_classes[l], np.float64)
        p = 1.0 / (1.0 + np.exp(-x))
        p_t = y * p + (1 - y) * (1 - p)
        f = (y * alpha + (1 - y) * (1 - alpha)) * np.power(1 - p_t, gamma) * (np.maximum(x, 0) - x * y + np.log1p(np.exp(-np.abs(x))))
        t = np.asarray(y_true_boxes[l], np.float64)
        e = np.abs(np.asarray(y_pred_boxes[l], np.float64) - t)
        hub = np.where(e <= delta, 0.5 * e * e, delta * e - 0.5 * delta * delta) * (t != 0)
        total += 50.0 * hub.sum() / (4.0 * npos) + f.sum() / npos / x.size
    return total
